"""`switch.closed` cells of line-bus / trafo-bus switches and LV-side tap changers as per-environment cells
(reference: opfgym/examples/network_reconfiguration.py:34-35, 49-60 -- switches and taps as actions;
security_constrained.py:31 -- 'closed' as contingency column).  pandapower hangs a line that is open at one
end from an auxiliary bus; the engine eliminates that bus analytically inside the fixed pattern.  The oracle
rebuilds its ppc (with auxiliary buses, oracle/ppc_ref.py) from the changed net for every environment."""
import numpy as np
import pytest
import torch

from opfgym_b200 import grids
from opfgym_b200 import net as pn
from opfgym_b200.opf_env import BatchedOpfEnv
from oracle import pf, scoring
from tests.hostsim.harness import TorchHostSimEngine


def make_env(n, tap_side="lv", **kw):
    net, profiles = grids.build_simbench_net("1-MV-comm--2-sw", n_profile_steps=96, load_scaling=1.5,
                                             gen_scaling=1.2)
    regular = list(net.line.index[net.line.in_service.to_numpy(bool)])
    ties = list(net.line.index[~net.line.in_service.to_numpy(bool)])
    net.line["in_service"] = True                       # the ties are energised; their SWITCHES are open
    sw = []
    for l in ties:                                      # a tie: switch at the to end, open
        sw.append(pn.create_switch(net, int(net.line.to_bus.loc[l]), int(l), "l", closed=False))
    for l in (regular[3], regular[len(regular) // 2]):  # regular lines: one switch at each end
        sw.append(pn.create_switch(net, int(net.line.from_bus.loc[l]), int(l), "l", closed=True))
        sw.append(pn.create_switch(net, int(net.line.to_bus.loc[l]), int(l), "l", closed=True))
    t1 = int(net.trafo.index[1])                        # second transformer: switch on its LV side
    sw.append(pn.create_switch(net, int(net.trafo.lv_bus.loc[t1]), t1, "t", closed=True))
    net.switch["min_closed"] = 0.0
    net.switch["max_closed"] = 1.0
    net.trafo["tap_side"] = tap_side
    net.trafo["tap_pos"] = 1.0                          # static position off neutral: the ppc row is built there
    net.trafo["min_tap_pos"] = -3.0
    net.trafo["max_tap_pos"] = 3.0
    for idx in net.ext_grid.index:
        pn.create_poly_cost(net, idx, "ext_grid", cp1_eur_per_mw=1.0)
    obs_keys = [("sgen", "p_mw", net.sgen.index), ("load", "p_mw", net.load.index),
                ("load", "q_mvar", net.load.index)]
    act_keys = [("switch", "closed", np.array(sw)), ("trafo", "tap_pos", net.trafo.index)]
    env = BatchedOpfEnv(net, act_keys, obs_keys, profiles=profiles, num_envs=n, train_data="full_uniform",
                        test_data="full_uniform", seed=1, obs_dtype="float64", **kw)
    return env, np.array(sw)


def _check(tap_side, kw):
    n = 20
    env, sw = make_env(n, tap_side, **kw)
    ns = len(sw)
    env.reset(seed=3)
    act = torch.rand(n, ns + len(env.net.trafo), dtype=torch.float64, generator=torch.Generator().manual_seed(9))
    act[0, :ns] = 1.0                    # everything closed
    act[1, :ns] = 0.0                    # everything open: islands, one transformer gone
    act[2, :ns] = torch.tensor([0, 0, 0, 0, 1, 0, 0, 1, 1], dtype=torch.float64)[:ns]   # lines open at ONE end
    e = env.engine
    state_before = e.state.clone()
    e.actions.copy_(act.to(env.device))
    e.step()
    assert bool(e.converged.all())
    vm = e.vm.cpu().numpy()
    loading = env.col("res_line", "loading_percent").cpu().numpy().copy()
    t_loading = env.col("res_trafo", "loading_percent").cpu().numpy().copy()
    reward, valids = e.reward.cpu().numpy().copy(), e.valids.cpu().numpy().copy()
    e.state.copy_(state_before)
    lk = env.program.ppc.bus_lookup
    has = lk >= 0
    assert has.all()
    seen_half_open = 0
    for b in range(n):
        net = env.net.deepcopy()
        for t, c in (("load", "p_mw"), ("load", "q_mvar"), ("sgen", "p_mw")):
            net[t][c] = env.col(t, c)[b].cpu().numpy()
        a = np.clip(act[b].numpy(), 0, 1)
        net.switch.loc[sw, "closed"] = np.round(a[:ns]).astype(bool)      # opf_env.py:476-478
        net.trafo["tap_pos"] = np.round(a[ns:] * 6.0 - 3.0)               # :479-481
        pf.runpp(net)                 # fresh ppc: auxiliary buses for half-open lines, islands dropped
        out = scoring.step_reward(net, env.constraints, env.reward_function)
        want_vm = net.res_bus.vm_pu.to_numpy()
        got_vm = vm[b][lk]
        assert (np.isnan(got_vm) == np.isnan(want_vm)).all(), b
        np.testing.assert_allclose(got_vm[~np.isnan(want_vm)], want_vm[~np.isnan(want_vm)], atol=1e-9)
        for got, want in ((loading[b], net.res_line.loading_percent.to_numpy()),
                          (t_loading[b], net.res_trafo.loading_percent.to_numpy())):
            assert (np.isnan(got) == np.isnan(want)).all(), b
            np.testing.assert_allclose(got[~np.isnan(want)], want[~np.isnan(want)], atol=1e-6)
        closed = net.switch.closed.to_numpy(bool)
        seen_half_open += int(closed[-5] != closed[-4]) + int(closed[-3] != closed[-2])
        np.testing.assert_allclose(reward[b], out["reward"], rtol=1e-8, atol=1e-10)
        assert (valids[b, :len(env.constraints)].astype(bool) == out["valids"]).all()
    assert seen_half_open >= 5


@pytest.mark.parametrize("tap_side", ["lv", "hv"])
def test_switch_cells_and_taps_hostsim(tap_side):
    _check(tap_side, dict(engine_cls=TorchHostSimEngine))


@pytest.mark.gpu
@pytest.mark.parametrize("tap_side", ["lv", "hv"])
def test_switch_cells_and_taps_cuda(cuda_lib, tap_side):
    _check(tap_side, {})


def test_bus_bus_switch_actions_are_rejected():
    net, profiles = grids.build_simbench_net("1-MV-comm--2-sw", n_profile_steps=96)
    s = pn.create_switch(net, int(net.bus.index[3]), int(net.bus.index[4]), "b", closed=False)
    net.switch["min_closed"] = 0.0
    net.switch["max_closed"] = 1.0
    with pytest.raises(NotImplementedError):
        BatchedOpfEnv(net, [("switch", "closed", np.array([s]))], [("load", "p_mw", net.load.index)],
                      profiles=profiles, num_envs=2, engine_cls=TorchHostSimEngine)
