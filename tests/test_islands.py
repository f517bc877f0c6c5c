"""Island handling (VERDICT r1 missing item 7): an in-service cell that cuts buses off every
slack bus.  pandapower drops those buses (bus type NONE, NaN results) and solves the rest; so does the
oracle, which rebuilds its ppc from the changed net (oracle/ppc_ref.py connectivity walk).  The engine keeps
its static pattern: kernel 1 finds the dropped buses per environment, the power-flow kernels hold them at
V = 0 behind identity rows.  Checked here: the N-1 loop of opfgym/security_constrained.py:37-68 on a RADIAL
grid, where every line outage islands the feeder behind it, and in-service cells as actions."""
import numpy as np
import pytest
import torch

from opfgym_b200 import grids
from opfgym_b200 import net as pn
from opfgym_b200 import reward as R
from opfgym_b200.net import LoadflowNotConverged
from opfgym_b200.opf_env import BatchedOpfEnv
from opfgym_b200.security_constrained import SecurityConstrainedBatchedOpfEnv
from oracle import pf, scoring
from tests.hostsim.harness import TorchHostSimEngine


def make_n1_env(n, drop_ties=False, **kw):
    net, profiles = grids.build_simbench_net("1-MV-semiurb--1-sw", n_profile_steps=96)
    if drop_ties:      # without the open tie lines the pattern is radial: the fused radial kernel runs
        net.line = net.line[net.line.in_service.to_numpy(bool)]
    net.sgen["controllable"] = net.sgen.max_max_p_mw > np.sort(net.sgen.max_max_p_mw.to_numpy())[-9]
    net.sgen["min_p_mw"] = 0.0
    net.sgen["max_p_mw"] = net.sgen.max_max_p_mw
    for idx in net.sgen.index[net.sgen.controllable]:
        pn.create_poly_cost(net, idx, "sgen", cp1_eur_per_mw=-0.03)
    in_service = net.line.index[net.line.in_service.to_numpy(bool)]
    # a line next to the substation (a whole feeder goes dark), one mid-feeder, one at a feeder's end
    outages = np.array([in_service[0], in_service[len(in_service) // 2], in_service[-1]])
    obs_keys = [("load", "p_mw", net.load.index), ("sgen", "p_mw", net.sgen.index)]
    act_keys = [("sgen", "p_mw", net.sgen.index[net.sgen.controllable])]
    env = SecurityConstrainedBatchedOpfEnv(
        net, act_keys, obs_keys, profiles=profiles, num_envs=n, train_data="full_uniform",
        test_data="full_uniform", seed=3, obs_dtype="float64",
        n_minus_one_keys=[("line", "in_service", outages)], not_converged_penalty=2.0,
        reward_function=R.Summation(), **kw)
    return env, outages


def reference_loop(env, outages, b, action, table="line", column="in_service"):
    """security_constrained.py:37-68 on a single pandas net; ``pf.runpp`` rebuilds the ppc, so islands are
    dropped the way pandapower drops them."""
    net = env.net.deepcopy()
    for t, c in (("load", "p_mw"), ("sgen", "p_mw")):
        net[t][c] = env._state_before[t, c][b]
    idxs = env.act_keys[0][2]
    lo, hi = net.sgen.min_p_mw.loc[idxs].to_numpy(), net.sgen.max_p_mw.loc[idxs].to_numpy()
    net.sgen.loc[idxs, "p_mw"] = (np.clip(action, 0, 1) * (hi - lo) + lo) / net.sgen.scaling.loc[idxs].to_numpy()
    pf.runpp(net)
    base = scoring.step_reward(net, env.constraints, env.reward_function)
    valids, viol, pens = base["valids"].copy(), base["violations"].copy(), base["unscaled_penalties"].copy()
    dropped = []
    for idx in outages:
        if not net[table].at[idx, column]:      # security_constrained.py:46-48
            continue
        net[table].at[idx, column] = False
        try:
            pf.runpp(net)
            dropped.append(int(np.isnan(net.res_bus.vm_pu.to_numpy()).sum()))
            m = [scoring.violation_metrics(c, net) for c in env.constraints]
            valids &= np.array([x["valid"] for x in m])
            viol += np.array([x["violation"] for x in m])
            pens += np.array([x["penalty"] for x in m])
        except LoadflowNotConverged:
            valids[:] = False
            viol += env.not_converged_penalty
            pens += env.not_converged_penalty
        net[table].at[idx, column] = True
    reward = scoring.reward(env.reward_function, base["objective"], pens.sum(), bool(valids.all()))
    return valids, viol, pens, reward, dropped


def _check_n1(kw):
    n = 5
    expect_kernel = kw.pop("expect_kernel")
    env, outages = make_n1_env(n, **kw)
    assert env.engine.info["pf_kernel_used"] == expect_kernel
    env.reset(seed=7)
    env._state_before = {(t, c): env.col(t, c).cpu().numpy().copy()
                         for t, c in (("load", "p_mw"), ("sgen", "p_mw"))}
    act = torch.rand(n, env.single_action_space.shape[0], dtype=torch.float64,
                     generator=torch.Generator().manual_seed(4))
    obs, reward, term, trunc, info = env.step(act)
    assert info["converged"].all()
    for b in range(n):
        valids, viol, pens, r, dropped = reference_loop(env, outages, b, act[b].numpy())
        assert min(dropped) >= 1 and max(dropped) >= 5, dropped     # every outage islands something
        np.testing.assert_array_equal(info["valids"][b].cpu().numpy(), valids)
        # no contingency counted as a failed power flow (that would add not_converged_penalty = 2 per constraint)
        np.testing.assert_allclose(info["violations"][b].cpu().numpy(), viol, rtol=1e-7, atol=1e-9)
        np.testing.assert_allclose(info["unscaled_penalties"][b].cpu().numpy(), pens, rtol=1e-7, atol=1e-9)
        np.testing.assert_allclose(float(reward[b]), r, rtol=1e-7, atol=1e-9)


def test_n_minus_one_islands_hostsim_radial_kernel():
    _check_n1(dict(engine_cls=TorchHostSimEngine, drop_ties=True, expect_kernel=3))


def test_n_minus_one_islands_hostsim_block_kernel():
    _check_n1(dict(engine_cls=TorchHostSimEngine, expect_kernel=1))      # open ties in the pattern: meshed


@pytest.mark.gpu
@pytest.mark.parametrize("drop_ties,expect_kernel", [(True, 3), (False, 1)])
def test_n_minus_one_islands_cuda(cuda_lib, drop_ties, expect_kernel):
    _check_n1(dict(drop_ties=drop_ties, expect_kernel=expect_kernel))


# ------------------------------------------------------------------ in-service cells as actions
def make_switch_env(n, **kw):
    """Ties AND three regular lines of 1-MV-comm--2-sw carry an in-service action: opening a regular line
    islands the feeder behind it unless a closed tie feeds it from the other side."""
    net, profiles = grids.build_simbench_net("1-MV-comm--2-sw", n_profile_steps=96, load_scaling=1.5,
                                             gen_scaling=1.2)
    net.line["min_in_service"] = 0.0
    net.line["max_in_service"] = 1.0
    for idx in net.ext_grid.index:
        pn.create_poly_cost(net, idx, "ext_grid", cp1_eur_per_mw=1.0)
    ties = list(net.line.index[~net.line.in_service.to_numpy(bool)])
    regular = list(net.line.index[net.line.in_service.to_numpy(bool)])
    switched = np.array([regular[2], regular[len(regular) // 2], regular[-2]] + ties)
    obs_keys = [("sgen", "p_mw", net.sgen.index), ("load", "p_mw", net.load.index),
                ("load", "q_mvar", net.load.index)]
    act_keys = [("line", "in_service", switched)]
    env = BatchedOpfEnv(net, act_keys, obs_keys, profiles=profiles, num_envs=n, train_data="full_uniform",
                        test_data="full_uniform", seed=1, obs_dtype="float64", **kw)
    return env, switched


def _check_switch(kw):
    n = 16
    env, switched = make_switch_env(n, **kw)
    env.reset(seed=3)
    act = torch.rand(n, len(switched), dtype=torch.float64, generator=torch.Generator().manual_seed(5))
    act[0] = 1.0                      # everything closed: meshed, nothing dropped
    act[1] = 0.0                      # everything open: three islands
    act[2, :3] = 0.0; act[2, 3:] = 1.0    # regular lines open, every tie closed: fed from the other side
    e = env.engine
    state_before = e.state.clone()
    e.actions.copy_(act.to(env.device))
    e.step()
    assert bool(e.converged.all())
    vm = e.vm.cpu().numpy()
    n_dropped = np.isnan(vm).sum(axis=1)
    assert n_dropped[0] == 0 and n_dropped[1] >= 3 and len(set(n_dropped.tolist())) >= 3
    loading = env.col("res_line", "loading_percent").cpu().numpy().copy()
    reward, valids = e.reward.cpu().numpy().copy(), e.valids.cpu().numpy().copy()
    e.state.copy_(state_before)
    lk = env.program.ppc.bus_lookup
    has = lk >= 0
    for b in range(n):
        net = env.net.deepcopy()
        for t, c in (("load", "p_mw"), ("load", "q_mvar"), ("sgen", "p_mw")):
            net[t][c] = env.col(t, c)[b].cpu().numpy()
        net.line.loc[switched, "in_service"] = np.round(np.clip(act[b].numpy(), 0, 1)).astype(bool)
        pf.runpp(net)                 # fresh ppc: the oracle's own connectivity walk drops the islands
        out = scoring.step_reward(net, env.constraints, env.reward_function)
        want_vm = net.res_bus.vm_pu.to_numpy()[has]
        got_vm = vm[b][lk[has]]
        assert (np.isnan(got_vm) == np.isnan(want_vm)).all(), (b, np.isnan(got_vm).sum(), np.isnan(want_vm).sum())
        np.testing.assert_allclose(got_vm[~np.isnan(want_vm)], want_vm[~np.isnan(want_vm)], atol=1e-9)
        want = net.res_line.loading_percent.to_numpy()
        assert (np.isnan(loading[b]) == np.isnan(want)).all()
        np.testing.assert_allclose(loading[b][~np.isnan(want)], want[~np.isnan(want)], atol=1e-6)
        np.testing.assert_allclose(reward[b], out["reward"], rtol=1e-8, atol=1e-10)
        assert (valids[b, :len(env.constraints)].astype(bool) == out["valids"]).all()


def test_switch_actions_with_islands_hostsim():
    _check_switch(dict(engine_cls=TorchHostSimEngine))


@pytest.mark.gpu
def test_switch_actions_with_islands_cuda(cuda_lib):
    _check_switch({})


# ------------------------------------------------------------------ 'closed' as contingency column
def _check_switch_contingencies(kw):
    """security_constrained.py:31 allows column 'closed': switches of line-bus type as N-1 elements."""
    n = 4
    net, profiles = grids.build_simbench_net("1-MV-semiurb--1-sw", n_profile_steps=96)
    net.sgen["controllable"] = net.sgen.max_max_p_mw > np.sort(net.sgen.max_max_p_mw.to_numpy())[-9]
    net.sgen["min_p_mw"] = 0.0
    net.sgen["max_p_mw"] = net.sgen.max_max_p_mw
    for idx in net.sgen.index[net.sgen.controllable]:
        pn.create_poly_cost(net, idx, "sgen", cp1_eur_per_mw=-0.03)
    lines = net.line.index[net.line.in_service.to_numpy(bool)]
    sw = [pn.create_switch(net, int(net.line.from_bus.loc[lines[5]]), int(lines[5]), "l", closed=True),
          pn.create_switch(net, int(net.line.to_bus.loc[lines[40]]), int(lines[40]), "l", closed=True),
          pn.create_switch(net, int(net.line.to_bus.loc[lines[70]]), int(lines[70]), "l", closed=False)]
    obs_keys = [("load", "p_mw", net.load.index), ("sgen", "p_mw", net.sgen.index)]
    act_keys = [("sgen", "p_mw", net.sgen.index[net.sgen.controllable])]
    env = SecurityConstrainedBatchedOpfEnv(
        net, act_keys, obs_keys, profiles=profiles, num_envs=n, train_data="full_uniform",
        test_data="full_uniform", seed=3, obs_dtype="float64",
        n_minus_one_keys=[("switch", "closed", np.array(sw))], not_converged_penalty=2.0,
        reward_function=R.Summation(), **kw)
    env.reset(seed=7)
    env._state_before = {(t, c): env.col(t, c).cpu().numpy().copy() for t, c in (("load", "p_mw"), ("sgen", "p_mw"))}
    act = torch.rand(n, env.single_action_space.shape[0], dtype=torch.float64,
                     generator=torch.Generator().manual_seed(4))
    obs, reward, term, trunc, info = env.step(act)
    assert info["converged"].all()
    for b in range(n):
        valids, viol, pens, r, dropped = reference_loop(env, sw, b, act[b].numpy(), "switch", "closed")
        assert len(dropped) == 2 and min(dropped) >= 1       # the switch that is open already is skipped
        np.testing.assert_array_equal(info["valids"][b].cpu().numpy(), valids)
        np.testing.assert_allclose(info["violations"][b].cpu().numpy(), viol, rtol=1e-7, atol=1e-9)
        np.testing.assert_allclose(info["unscaled_penalties"][b].cpu().numpy(), pens, rtol=1e-7, atol=1e-9)
        np.testing.assert_allclose(float(reward[b]), r, rtol=1e-7, atol=1e-9)


def test_switch_contingencies_hostsim():
    _check_switch_contingencies(dict(engine_cls=TorchHostSimEngine))


@pytest.mark.gpu
def test_switch_contingencies_cuda(cuda_lib):
    _check_switch_contingencies({})


def test_ties_are_not_island_critical():
    """Only a switchable branch that is an edge of the static grid's spanning forest can cut buses off; the
    forest is grown through normally-closed branches first, so normally-open ties never trigger kernel 1's
    connectivity walk (with ties as forest edges BASELINE config 5 lost 20 % of its throughput)."""
    from tests.test_dynamic_branches import make_env
    env, ties = make_env(3, engine_cls=TorchHostSimEngine)
    n_lines_in_service = int(env.net.line.in_service.to_numpy(bool).sum())
    assert env.engine.bry.shape[1] == len(env.net.line) + len(env.net.trafo)         # every branch has a cell
    # the whole line column is per-environment: exactly the regular lines are critical, none of the four ties
    assert env.engine.info["n_island_critical"] == n_lines_in_service
    env, switched = make_switch_env(3, engine_cls=TorchHostSimEngine)
    assert env.engine.info["n_island_critical"] == n_lines_in_service
