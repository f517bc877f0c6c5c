"""N-1 security-constrained env against a per-environment replay of the
reference's loop (opfgym/security_constrained.py:37-68) on the CPU oracle."""
import networkx as nx
import numpy as np
import pytest
import torch

from opfgym_b200 import grids
from opfgym_b200 import net as pn
from opfgym_b200 import reward as R
from opfgym_b200.net import LoadflowNotConverged
from opfgym_b200.ppc import PpcBuilder
from opfgym_b200.security_constrained import SecurityConstrainedBatchedOpfEnv
from oracle import pf, scoring
from tests.hostsim.harness import TorchHostSimEngine


def make_env(n, **kw):
    net, profiles = grids.build_simbench_net("1-HV-mixed--1-sw", n_profile_steps=96, load_scaling=0.8,
                                             gen_scaling=0.8, max_loading=30)
    net.sgen["controllable"] = net.sgen.max_max_p_mw > 24
    net.sgen["min_p_mw"] = 0.0
    net.sgen["max_p_mw"] = net.sgen.max_max_p_mw
    for idx in net.sgen.index:
        pn.create_poly_cost(net, idx, "sgen", cp1_eur_per_mw=-0.03)
    g = nx.MultiGraph()
    g.add_edges_from(zip(net.line.from_bus, net.line.to_bus, net.line.index))
    bridges = {frozenset(e) for e in nx.bridges(nx.Graph(g))}
    loop_lines = [i for i, (f, t) in enumerate(zip(net.line.from_bus, net.line.to_bus))
                  if frozenset((f, t)) not in bridges]
    outages = np.array(loop_lines[:2] + loop_lines[40:41])
    obs_keys = [("load", "p_mw", net.load.index), ("sgen", "p_mw", net.sgen.index)]
    act_keys = [("sgen", "p_mw", net.sgen.index[net.sgen.controllable])]
    env = SecurityConstrainedBatchedOpfEnv(
        net, act_keys, obs_keys, profiles=profiles, num_envs=n, train_data="full_uniform",
        test_data="full_uniform", seed=2, obs_dtype="float64",
        n_minus_one_keys=[("line", "in_service", outages)], not_converged_penalty=2.0,
        reward_function=R.Parameterized(valid_reward=0.5, invalid_penalty=0.25, penalty_weight=0.4), **kw)
    return env, outages


def reference_loop(env, outages, b, action):
    """security_constrained.py:37-68 on a single pandas net."""
    net = env.net.deepcopy()
    for t, c in (("load", "p_mw"), ("sgen", "p_mw")):
        net[t][c] = env._state_before[t, c][b]
    idxs = env.act_keys[0][2]
    lo, hi = net.sgen.min_p_mw.loc[idxs].to_numpy(), net.sgen.max_p_mw.loc[idxs].to_numpy()
    net.sgen.loc[idxs, "p_mw"] = (np.clip(action, 0, 1) * (hi - lo) + lo) / net.sgen.scaling.loc[idxs].to_numpy()
    pf.runpp(net)
    base = scoring.step_reward(net, env.constraints, env.reward_function)
    valids, viol, pens = base["valids"].copy(), base["violations"].copy(), base["unscaled_penalties"].copy()
    for idx in outages:
        if not net.line.at[idx, "in_service"]:
            continue
        net.line.at[idx, "in_service"] = False
        try:
            pf.runpp(net)
            m = [scoring.violation_metrics(c, net) for c in env.constraints]
            valids &= np.array([x["valid"] for x in m])
            viol += np.array([x["violation"] for x in m])
            pens += np.array([x["penalty"] for x in m])
        except LoadflowNotConverged:
            valids[:] = False
            viol += env.not_converged_penalty
            pens += env.not_converged_penalty
        net.line.at[idx, "in_service"] = True
    reward = scoring.reward(env.reward_function, base["objective"], pens.sum(), bool(valids.all()))
    return valids, viol, pens, reward


def _check(kw):
    n = 6
    env, outages = make_env(n, **kw)
    env.reset(seed=5)
    env._state_before = {(t, c): env.col(t, c).cpu().numpy().copy()
                         for t, c in (("load", "p_mw"), ("sgen", "p_mw"))}
    act = torch.rand(n, env.single_action_space.shape[0], dtype=torch.float64,
                     generator=torch.Generator().manual_seed(1))
    obs, reward, term, trunc, info = env.step(act)
    assert info["converged"].all() and term.all()
    worst_valid_drop = 0
    for b in range(n):
        valids, viol, pens, r = reference_loop(env, outages, b, act[b].numpy())
        np.testing.assert_array_equal(info["valids"][b].cpu().numpy(), valids)
        np.testing.assert_allclose(info["violations"][b].cpu().numpy(), viol, rtol=1e-7, atol=1e-9)
        np.testing.assert_allclose(info["unscaled_penalties"][b].cpu().numpy(), pens, rtol=1e-7, atol=1e-9)
        np.testing.assert_allclose(float(reward[b]), r, rtol=1e-7, atol=1e-9)
    assert (info["violations"] > 0).any()       # the contingencies do bite (max_loading=30)


@pytest.mark.parametrize("batched", [True, False])
def test_n_minus_one_hostsim(batched):
    """batched: all (environment, contingency) pairs as ONE batch of rows; else one pass per contingency."""
    _check(dict(engine_cls=TorchHostSimEngine, batch_contingencies=batched))


@pytest.mark.gpu
@pytest.mark.parametrize("batched", [True, False])
def test_n_minus_one_cuda(cuda_lib, batched):
    _check(dict(batch_contingencies=batched))
