"""``net.ward`` and ``net.impedance`` in the in-repo net -> ppc conversion (element
coverage of the conversion; pandapower build_bus ``_calc_pq_elements_and_add_on_ppc`` /
``_calc_shunts_and_add_on_ppc`` and build_branch ``_calc_impedance_parameter``): the product's builder
against the oracle's own, and the engine's power flow against the oracle's per environment -- ward
constant-power part as a per-environment ACTION cell, impedance elements closing a mesh."""
import numpy as np
import pytest
import torch

from opfgym_b200 import grids, net as N, ppc as P
from opfgym_b200.opf_env import BatchedOpfEnv
from oracle import pf, ppc_ref as R
from tests.hostsim.harness import TorchHostSimEngine


def _net(name):
    net, profiles = grids.build_simbench_net(name, n_profile_steps=96)
    buses = net.bus.index.to_numpy()
    live = buses[P.PpcBuilder(net).build(net).bus_lookup >= 0]
    rng = np.random.default_rng(7)
    pick = rng.choice(live[1:], 6, replace=False)
    N.create_ward(net, pick[0], ps_mw=0.4, qs_mvar=0.1, pz_mw=0.2, qz_mvar=-0.15)
    N.create_ward(net, pick[1], ps_mw=-0.3, qs_mvar=0.05, pz_mw=0.0, qz_mvar=0.3)
    N.create_ward(net, pick[2], ps_mw=9.0, qs_mvar=9.0, pz_mw=9.0, qz_mvar=9.0, in_service=False)
    # same voltage level only: an impedance element has no ratio
    vn = net.bus.vn_kv
    same = [b for b in live if vn[b] == vn[pick[3]] and b != pick[3]]
    N.create_impedance(net, pick[3], same[0], rft_pu=0.02, xft_pu=0.06, sn_mva=10.0)
    N.create_impedance(net, pick[3], same[-1], rft_pu=0.01, xft_pu=0.03, sn_mva=net.sn_mva, in_service=False)
    net.ward["min_ps_mw"], net.ward["max_ps_mw"] = -0.5, 0.5
    return net, profiles


@pytest.mark.parametrize("name", ["1-MV-rural--0-sw", "1-HV-urban--0-sw"])
def test_builders_agree_on_ward_and_impedance(name):
    net, _ = _net(name)
    a, b = P.PpcBuilder(net).build(net), R.build(net)
    assert a.bus.shape == b.bus.shape and a.branch.shape == b.branch.shape
    ok = a.bus_lookup >= 0
    perm = np.full(a.bus.shape[0], -1)
    perm[a.bus_lookup[ok]] = b.bus_lookup[ok]
    for la, lb in zip(a.line_branch, b.line_branch):
        if la >= 0:
            for col in (P.F_BUS, P.T_BUS):
                perm[int(a.branch[la, col])] = int(b.branch[lb, col])
    for col in (P.PD, P.QD, P.GS, P.BS):
        np.testing.assert_allclose(a.bus[:, col], b.bus[perm, col], rtol=1e-12, atol=1e-12)
    assert np.abs(a.bus[:, P.GS]).sum() > 0 and np.abs(a.bus[:, P.BS]).sum() > 0
    assert list(a.impedance_branch >= 0) == [True, False] == list(b.impedance_branch >= 0)
    ra, rb = int(a.impedance_branch[0]), int(b.impedance_branch[0])
    assert ra == a.branch.shape[0] - 1 and rb == b.branch.shape[0] - 1      # stacked behind lines and trafos
    np.testing.assert_allclose(a.branch[ra, 2:], b.branch[rb, 2:], rtol=1e-13)
    assert a.branch[ra, P.BR_R] == pytest.approx(0.02 * net.sn_mva / 10.0)
    assert a.rate_f[ra] == 0.0 and a.rate_t[ra] == 0.0


def test_asymmetric_impedance_is_rejected():
    net, _ = _net("1-MV-rural--0-sw")
    net.impedance.loc[net.impedance.index[0], "xtf_pu"] = 0.09
    with pytest.raises(NotImplementedError, match="symmetric"):
        P.PpcBuilder(net).build(net)


def _check(name, sync=lambda: None, **kw):
    net, profiles = _net(name)
    n = 5
    obs_keys = [("load", "p_mw", net.load.index), ("res_bus", "vm_pu", net.bus.index[:6])]
    act_keys = [("ward", "ps_mw", net.ward.index[:2])]
    env = BatchedOpfEnv(net, act_keys, obs_keys, profiles=profiles, num_envs=n, train_data="full_uniform",
                        test_data="full_uniform", seed=5, obs_dtype="float64", **kw)
    env.reset(seed=6)
    act = torch.rand(n, 2, dtype=torch.float64, generator=torch.Generator().manual_seed(8))
    e = env.engine
    e.actions.copy_(act.to(env.device))
    state = e.state.clone()
    e.step()
    sync()
    assert bool(e.converged.all())
    lk = env.program.ppc.bus_lookup
    for b in range(n):
        one = env.net.deepcopy()
        for t, c in (("load", "p_mw"), ("load", "q_mvar"), ("sgen", "p_mw"), ("storage", "p_mw")):
            if env.program.layout.has(t, c):
                one[t][c] = state[b, env.program.layout.slice(t, c)].cpu().numpy()
        ps = one.ward.ps_mw.to_numpy().copy()
        ps[:2] = -0.5 + act[b].numpy()
        one.ward["ps_mw"] = ps
        pf.runpp(one)                                            # oracle conversion + oracle solver
        live = lk >= 0
        np.testing.assert_allclose(e.vm[b].cpu().numpy()[lk[live]], one.res_bus.vm_pu.to_numpy()[live], atol=1e-9)
        np.testing.assert_allclose(np.degrees(e.va[b].cpu().numpy()[lk[live]]),
                                   one.res_bus.va_degree.to_numpy()[live], atol=1e-7)
    # the elements matter: without them the solution moves
    bare = env.net.deepcopy()
    bare.ward.drop(bare.ward.index, inplace=True)
    bare.impedance.drop(bare.impedance.index, inplace=True)
    pf.runpp(bare)
    pf.runpp(one := env.net.deepcopy())
    assert np.nanmax(np.abs(bare.res_bus.vm_pu.to_numpy() - one.res_bus.vm_pu.to_numpy())) > 1e-5


@pytest.mark.parametrize("name", ["1-MV-rural--0-sw", "1-HV-urban--0-sw"])
def test_engine_solves_ward_and_impedance_hostsim(name):
    _check(name, engine_cls=TorchHostSimEngine)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["1-MV-rural--0-sw", "1-HV-urban--0-sw"])
def test_engine_solves_ward_and_impedance_cuda(cuda_lib, name):
    _check(name, sync=torch.cuda.synchronize)
