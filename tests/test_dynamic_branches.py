"""BASELINE config 5: discrete tap / line in-service actions -> per-environment Ybus values
(kernel 1 rebuilds branch admittances and Ybus per env in the fixed pattern).  Action semantics:
rounding rules of opfgym/opf_env.py:476-481 as used by examples/network_reconfiguration.py:34-35."""
import numpy as np
import pytest
import torch

from opfgym_b200 import grids
from opfgym_b200 import net as pn
from opfgym_b200.opf_env import BatchedOpfEnv
from opfgym_b200.ppc import PpcBuilder
from oracle import pf, scoring
from tests.hostsim.harness import TorchHostSimEngine


def make_env(n, **kw):
    net, profiles = grids.build_simbench_net("1-MV-comm--2-sw", n_profile_steps=96, load_scaling=2.2,
                                             gen_scaling=1.6)
    net.load["controllable"] = net.load.max_max_p_mw > 0.6
    net.load["min_p_mw"] = 0.0
    net.load["max_p_mw"] = net.load.max_max_p_mw
    net.trafo["min_tap_pos"] = -3.0
    net.trafo["max_tap_pos"] = 3.0
    net.line["min_in_service"] = 0.0
    net.line["max_in_service"] = 1.0
    net.ext_grid["max_p_mw"] = 8.0
    net.ext_grid["min_p_mw"] = -np.inf
    for idx in net.load.index[net.load.controllable]:
        pn.create_poly_cost(net, idx, "load", cp1_eur_per_mw=-3.0)
    ties = net.line.index[~net.line.in_service]
    assert len(ties) == 4
    obs_keys = [("sgen", "p_mw", net.sgen.index), ("load", "p_mw", net.load.index),
                ("load", "q_mvar", net.load.index)]
    act_keys = [("load", "p_mw", net.load.index[net.load.controllable]),
                ("trafo", "tap_pos", net.trafo.index), ("line", "in_service", ties)]
    env = BatchedOpfEnv(net, act_keys, obs_keys, profiles=profiles, num_envs=n,
                        train_data="full_uniform", test_data="full_uniform", seed=1,
                        obs_dtype="float64", **kw)
    return env, ties


def oracle(env, ties, b, action):
    net = env.net.deepcopy()
    for t, c in (("load", "p_mw"), ("load", "q_mvar"), ("sgen", "p_mw")):
        net[t][c] = env.col(t, c)[b].cpu().numpy()
    a = np.clip(action, 0, 1)
    k = 0
    for table, column, idxs in env.act_keys:
        df = net[table]
        lo, hi = df[f"min_{column}"].loc[idxs].to_numpy(float), df[f"max_{column}"].loc[idxs].to_numpy(float)
        sp = a[k:k + len(idxs)] * (hi - lo) + lo
        if "scaling" in df.columns:
            sp = sp / df.scaling.loc[idxs].to_numpy()
        if column == "in_service":
            sp = np.round(sp).astype(bool)
        elif column == "tap_pos":
            sp = np.round(sp)
        net[table].loc[idxs, column] = sp
        k += len(idxs)
    res = pf.runpp(net)        # topology changed: fresh builder
    out = scoring.step_reward(net, env.constraints, env.reward_function)
    return net, res, out


def _check(engine_kw):
    n = 24
    env, ties = make_env(n, **engine_kw)
    assert env.engine.info["nb"] == 111 and env.engine.bry.shape[1] == len(env.net.line) + 2
    env.reset(seed=3)
    g = torch.Generator().manual_seed(0)
    act = torch.rand(n, env.single_action_space.shape[0], dtype=torch.float64, generator=g)
    act[0, -4:] = 0.0          # all ties open, taps anywhere
    act[1, -4:] = 1.0          # all ties closed
    act[2, -6:-4] = 0.5        # neutral taps
    e = env.engine
    state_before = e.state.clone()
    e.actions.copy_(act.to(env.device))
    e.step()
    assert bool(e.converged.all())
    taps = env.col("trafo", "tap_pos").cpu().numpy()
    assert set(np.unique(taps)) <= set(np.arange(-3.0, 4.0)) and len(np.unique(taps)) > 3
    svc = env.col("line", "in_service").cpu().numpy()
    assert set(np.unique(svc)) == {0.0, 1.0}
    e.state.copy_(state_before)                 # oracle replays from the pre-action cells
    for b in range(n):
        net, res, out = oracle(env, ties, b, act[b].numpy())
        lk = env.program.ppc.bus_lookup          # pandapower bus order on both sides
        has = lk >= 0
        np.testing.assert_allclose(e.vm[b].cpu().numpy()[lk[has]], net.res_bus.vm_pu.to_numpy()[has], atol=1e-9)
        np.testing.assert_allclose(e.va[b].cpu().numpy()[lk[has]],
                                   np.radians(net.res_bus.va_degree.to_numpy()[has]), atol=1e-9)
    e.actions.copy_(act.to(env.device))
    e.step()
    for b in range(n):
        net, res, out = oracle(env, ties, b, act[b].numpy())
        got = env.col("res_line", "loading_percent")[b].cpu().numpy()
        want = net.res_line.loading_percent.to_numpy()
        assert (np.isnan(got) == np.isnan(want)).all()          # open ties report NaN like pandapower
        np.testing.assert_allclose(got[~np.isnan(want)], want[~np.isnan(want)], atol=1e-6)
        np.testing.assert_allclose(float(e.reward[b]), out["reward"], rtol=1e-8, atol=1e-10)
        assert (e.valids[b, :len(env.constraints)].cpu().numpy().astype(bool) == out["valids"]).all()
        e.state.copy_(state_before) if False else None


def test_tap_and_switch_actions_hostsim():
    _check(dict(engine_cls=TorchHostSimEngine))


@pytest.mark.gpu
def test_tap_and_switch_actions_cuda(cuda_lib):
    _check({})
