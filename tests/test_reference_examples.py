"""The reference's ``opfgym/examples`` written against this package (same structure: a subclass that defines the
OPF on the net, lists observation / action keys and overrides ``_sampling``), stepped as a batch and checked per
environment against the oracle's power flow and the reference's reward arithmetic:

* ``examples/mixed_continuous_discrete.py`` -- continuous (``sgen.q_mvar``) and discrete (``trafo.tap_pos``) actuators,
  a sampled slack voltage, and an objective that is not a pandapower cost (``objective_function=``, batched);
* ``examples/partial_obs.py``               -- observation keys are a subset of the state keys;
* ``examples/pure_constraint_satisfaction.py`` -- no objective at all, only constraints.

(``custom_constraint.py``: tests/test_custom_constraints.py; ``multi_stage.py``: test_multi_stage.py /
test_golden_f3.py; ``security_constrained.py``: test_security_constrained.py; ``network_reconfiguration.py``:
test_switch_cells.py / test_islands.py; ``stochastic_obs.py``: test_wrappers_mixed.py; ``non_simbench_net.py``:
test_ieee14_net.py.)"""
import numpy as np
import pytest
import torch

from opfgym_b200 import grids, net as pn
from opfgym_b200.opf_env import BatchedOpfEnv
from oracle import pf, scoring
from tests.hostsim.harness import TorchHostSimEngine

DATA = dict(train_data="full_uniform", test_data="full_uniform", obs_dtype="float64")


class MixedContinuousDiscrete(BatchedOpfEnv):
    def __init__(self, cos_phi=0.95, **kwargs):
        net, profiles = grids.build_simbench_net("1-MV-semiurb--1-sw", n_profile_steps=96)
        net.trafo["controllable"] = True
        net.trafo["min_tap_pos"], net.trafo["max_tap_pos"] = -2, 2
        net.sgen["controllable"] = True
        s_max = net.sgen.max_max_p_mw / cos_phi
        net.sgen["max_max_q_mvar"] = (s_max ** 2 - net.sgen.max_max_p_mw ** 2) ** 0.5
        net.sgen["min_min_q_mvar"] = -net.sgen.max_max_q_mvar
        net.sgen["max_q_mvar"], net.sgen["min_q_mvar"] = net.sgen.max_max_q_mvar, -net.sgen.max_max_q_mvar
        net.ext_grid["min_vm_pu"], net.ext_grid["max_vm_pu"] = 0.95, 1.05
        obs_keys = [("ext_grid", "vm_pu", net.ext_grid.index), ("sgen", "p_mw", net.sgen.index),
                    ("load", "p_mw", net.load.index), ("load", "q_mvar", net.load.index)]
        act_keys = [("sgen", "q_mvar", net.sgen.index), ("trafo", "tap_pos", net.trafo.index)]
        super().__init__(net, act_keys, obs_keys, profiles=profiles,
                         objective_function=lambda env: (env.col("res_bus", "vm_pu") - 1.0) ** 2, **kwargs)

    def _sampling(self, *args, **kwargs):
        super()._sampling(*args, **kwargs)
        self._sample_from_range("ext_grid", "vm_pu", self.net.ext_grid.index)      # the example's harder variant


class PartiallyObservable(BatchedOpfEnv):
    def __init__(self, observable_loads=np.arange(10), **kwargs):
        net, profiles = grids.build_simbench_net("1-MV-rural--0-sw", n_profile_steps=96)
        net.sgen["controllable"] = True
        net.sgen["min_p_mw"], net.sgen["max_p_mw"] = 0.0, net.sgen.max_max_p_mw
        net.sgen["min_q_mvar"] = net.sgen["max_q_mvar"] = 0.0
        for idx in net.ext_grid.index:
            pn.create_poly_cost(net, idx, "ext_grid", cp1_eur_per_mw=1)
        obs_keys = [("load", "p_mw", net.load.index[observable_loads]), ("load", "q_mvar", net.load.index[observable_loads])]
        state_keys = [("load", "p_mw", net.load.index), ("load", "q_mvar", net.load.index)]
        super().__init__(net, [("sgen", "p_mw", net.sgen.index)], obs_keys, state_keys=state_keys, profiles=profiles, **kwargs)


class ConstraintSatisfaction(BatchedOpfEnv):
    def __init__(self, **kwargs):
        net, profiles = grids.build_simbench_net("1-MV-rural--0-sw", n_profile_steps=96)
        net.sgen["controllable"] = True
        net.sgen["min_p_mw"], net.sgen["max_p_mw"] = 0.0, net.sgen.max_max_p_mw
        net.sgen["min_q_mvar"] = net.sgen["max_q_mvar"] = 0.0
        net.ext_grid["max_p_mw"] = 1.0
        net.bus["max_vm_pu"], net.bus["min_vm_pu"] = 1.02, 0.98
        net.line["max_loading_percent"] = 60.0
        obs_keys = [("load", "p_mw", net.load.index), ("load", "q_mvar", net.load.index)]
        super().__init__(net, [("sgen", "p_mw", net.sgen.index)], obs_keys, profiles=profiles, **kwargs)


def _oracle_net(env, state, actions_physical, b):
    """One pandapower-style net carrying environment b's cells, solved by the oracle."""
    net = env.net.deepcopy()
    lay = env.program.layout
    for (t, c) in list(lay.columns):
        if t.startswith("res_") or t in ("poly_cost", "pwl_cost") or c in ("closed",):
            continue
        if c in net[t].columns:
            net[t][c] = state[b, lay.slice(t, c)].cpu().numpy()
    for (t, c), vals in actions_physical.items():        # action columns cover the whole table here
        net[t][c] = np.asarray(vals[b], float)
    pf.runpp(net, enforce_q_lims=True)
    return net


def _step(env, seed, sync):
    env.reset(seed=seed)
    n_act = env.single_action_space.shape[0]
    act = torch.rand(env.num_envs, n_act, dtype=torch.float64, generator=torch.Generator().manual_seed(seed))
    e = env.engine
    state = e.state.clone()
    out = env.step(act.to(env.device))
    sync()
    assert bool(out[4]["converged"].all())
    return act, state, e, out


def _check_mixed(sync=lambda: None, **kw):
    env = MixedContinuousDiscrete(num_envs=5, seed=1, **DATA, **kw)
    assert env.single_action_space.shape[0] == len(env.net.sgen) + len(env.net.trafo)
    act, state, e, (obs, reward, term, trunc, info) = _step(env, 4, sync)
    ns = len(env.net.sgen)
    lo, hi = env.net.sgen.min_q_mvar.to_numpy(float), env.net.sgen.max_q_mvar.to_numpy(float)
    q = lo + act[:, :ns].numpy() * (hi - lo)
    tap = np.rint(-2 + act[:, ns:].numpy() * 4)                       # discrete actuator: rounded (opf_env.py:476-478)
    vm_slack = state[:, env.program.layout.slice("ext_grid", "vm_pu")].cpu().numpy()
    assert np.ptp(vm_slack) > 0.01 and (vm_slack >= 0.95).all() and (vm_slack <= 1.05).all()   # sampled per environment
    lk = env.program.ppc.bus_lookup
    for b in range(env.num_envs):
        net = _oracle_net(env, state, {("sgen", "q_mvar"): q, ("trafo", "tap_pos"): tap}, b)
        vm = net.res_bus.vm_pu.to_numpy()
        np.testing.assert_allclose(e.vm[b].cpu().numpy()[lk], vm, atol=1e-9)
        metrics = [scoring.violation_metrics(c, net) for c in env.constraints]
        pen, valid = sum(m["penalty"] for m in metrics), all(m["valid"] for m in metrics)
        want = scoring.reward(env.reward_function, -float(((vm - 1.0) ** 2).sum()), pen, valid)
        assert float(reward[b]) == pytest.approx(want, rel=1e-8, abs=1e-10)


def _check_partial(sync=lambda: None, **kw):
    env = PartiallyObservable(num_envs=4, seed=1, **DATA, **kw)
    assert env.single_observation_space.shape[0] == 20                 # ten loads, p and q
    obs0, _ = env.reset(seed=2)
    state = env.engine.state
    lay = env.program.layout
    np.testing.assert_array_equal(obs0[:, :10].cpu().numpy(), state[:, lay.slice("load", "p_mw")][:, :10].cpu().numpy())
    # the unobserved loads are sampled too (state keys) -- they differ between environments
    assert float(state[:, lay.slice("load", "p_mw")][:, 10:].std(dim=0).min()) > 0
    act, state, e, (obs, reward, *_rest) = _step(env, 2, sync)
    p = act.numpy() * env.net.sgen.max_max_p_mw.to_numpy(float)
    for b in range(env.num_envs):
        net = _oracle_net(env, state, {("sgen", "p_mw"): p / env.net.sgen.scaling.to_numpy(float)}, b)
        r = scoring.step_reward(net, env.constraints, env.reward_function)
        assert float(reward[b]) == pytest.approx(r["reward"], rel=1e-8, abs=1e-10)


def _check_constraints_only(sync=lambda: None, **kw):
    env = ConstraintSatisfaction(num_envs=6, seed=1, **DATA, **kw)
    kinds = sorted(type(c).__name__ for c in env.constraints)
    assert kinds == ["ExtGridActivePowerConstraint", "LineOverloadConstraint", "TrafoOverloadConstraint",
                     "VoltageConstraint"], kinds                    # the stand-in grid's transformer carries a limit too
    act, state, e, (obs, reward, term, trunc, info) = _step(env, 3, sync)
    assert float(e.objective.abs().max()) == 0.0                       # no objective: the reward is the penalty
    p = act.numpy() * env.net.sgen.max_max_p_mw.to_numpy(float)
    some_invalid = False
    for b in range(env.num_envs):
        net = _oracle_net(env, state, {("sgen", "p_mw"): p / env.net.sgen.scaling.to_numpy(float)}, b)
        r = scoring.step_reward(net, env.constraints, env.reward_function)
        assert float(reward[b]) == pytest.approx(r["reward"], rel=1e-8, abs=1e-10)
        assert info["valids"][b].cpu().numpy().tolist() == r["valids"].tolist()
        some_invalid |= not r["valid"]
    assert some_invalid                                                # the tightened limits bind


@pytest.mark.parametrize("check", [_check_mixed, _check_partial, _check_constraints_only])
def test_reference_examples_hostsim(check):
    check(engine_cls=TorchHostSimEngine)


@pytest.mark.gpu
@pytest.mark.parametrize("check", [_check_mixed, _check_partial, _check_constraints_only])
def test_reference_examples_cuda(cuda_lib, check):
    check(sync=torch.cuda.synchronize)
