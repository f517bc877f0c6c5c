"""IEEE 14-bus system as a NET of physical elements (lines in ohm / nF per km, transformers by vk / vkr / tap
changer, loads, a shunt, generators, an ext_grid) built with the in-repo pandapower-style constructors -- the
published MATPOWER solution then holds the whole chain  net -> ppc conversion -> power flow  against a published
answer, for the product's conversion + engine and for the oracle's own conversion + solver.  (tests/test_ieee14.py
enters at the ppc.)  What this pins: per-unit conversion of lines on their bus voltage, transformer impedance from
vk / vkr on the transformer's rating, the HV-side tap changer as an off-nominal ratio, the shunt's sign, PV set-points.

It is also the reference's ``examples/non_simbench_net.py`` (a standard case without time series: generator set-points
as actions, loads sampled ``normal_around_mean``) on this package."""
import numpy as np
import pytest
import torch

from opfgym_b200 import adapter, net as N
from opfgym_b200.opf_env import BatchedOpfEnv
from oracle import pf
from tests.hostsim.harness import TorchHostSimEngine
from tests.test_ieee14 import BRANCHES, GENS, LOADS, PG_SLACK, QG14, check_solution

F_HZ, SN = 50.0, 100.0
VN = {b: (135.0 if b <= 5 else 35.0) for b in range(1, 15)}        # the three tap transformers join the two levels


def ieee14_net():
    net = N.create_empty_network("ieee14", f_hz=F_HZ, sn_mva=SN)
    bus = {b: N.create_bus(net, vn_kv=VN[b], min_vm_pu=0.94, max_vm_pu=1.10) for b in range(1, 15)}
    for f, t, r, x, b, tap in BRANCHES:
        if tap:                                                    # MATPOWER: ratio at the from (HV) side, z at the to side
            N.create_transformer_from_parameters(
                net, bus[f], bus[t], sn_mva=SN, vn_hv_kv=VN[f], vn_lv_kv=VN[t], vkr_percent=100.0 * r,
                vk_percent=100.0 * np.hypot(r, x), pfe_kw=0.0, i0_percent=0.0, tap_side="hv", tap_neutral=0.0,
                tap_step_percent=0.1, tap_pos=(tap - 1.0) * 1000.0, tap_min=-100.0, tap_max=100.0)
        else:
            base_r = VN[f] ** 2 / SN
            N.create_line_from_parameters(net, bus[f], bus[t], length_km=1.0, r_ohm_per_km=r * base_r,
                                          x_ohm_per_km=x * base_r,
                                          c_nf_per_km=b / (2 * np.pi * F_HZ * 1e-9 * base_r), max_i_ka=1.0)
    for b, (p, q) in LOADS.items():
        N.create_load(net, bus[b], p_mw=p, q_mvar=q)
    N.create_shunt(net, bus[9], q_mvar=-19.0, vn_kv=VN[9])         # MATPOWER BS = +19 MVAr injected at 1 p.u.
    b0, _, v0, _, _ = GENS[0]
    N.create_ext_grid(net, bus[b0], vm_pu=v0)
    for b, pg, vg, qmax, qmin in GENS[1:]:
        N.create_gen(net, bus[b], p_mw=pg, vm_pu=vg)
    return net


def _check(net):
    check_solution(net.res_bus.vm_pu.to_numpy(), net.res_bus.va_degree.to_numpy())
    assert abs(float(net.res_ext_grid.p_mw.iloc[0]) - PG_SLACK) < 6e-3
    q = np.concatenate([net.res_ext_grid.q_mvar.to_numpy(), net.res_gen.q_mvar.to_numpy()])
    np.testing.assert_allclose(q, QG14, rtol=0, atol=6e-3)


def test_oracle_conversion_and_solver_reach_the_published_solution():
    net = ieee14_net()
    pf.runpp(net, tolerance_mva=1e-6)
    _check(net)


def test_product_conversion_and_engine_reach_the_published_solution():
    net = ieee14_net()
    adapter.PowerFlowSolver(net, engine_cls=TorchHostSimEngine, tolerance_mva=1e-6)(net)
    _check(net)


@pytest.mark.gpu
def test_product_conversion_and_cuda_engine_reach_the_published_solution(cuda_lib):
    net = ieee14_net()
    adapter.PowerFlowSolver(net, tolerance_mva=1e-6)(net)
    _check(net)


class NonSimbenchNet(BatchedOpfEnv):
    """examples/non_simbench_net.py: a standard case, loads ~ N(mean, 0.3 mean) truncated to +-30 %, generator
    active power as actions."""

    def __init__(self, **kwargs):
        net = ieee14_net()
        net.gen["min_p_mw"], net.gen["max_p_mw"] = 0.0, 80.0
        net.gen["controllable"] = True
        spread = 0.3
        for col in ("p_mw", "q_mvar"):
            v = net.load[col].to_numpy(float)
            net.load[f"min_min_{col}"], net.load[f"max_max_{col}"] = v - spread * np.abs(v), v + spread * np.abs(v)
            net.load[f"mean_{col}"], net.load[f"std_dev_{col}"] = v, spread * np.abs(v)
        N.create_poly_cost(net, 0, "ext_grid", cp1_eur_per_mw=1.0)
        obs_keys = [("load", "p_mw", net.load.index), ("load", "q_mvar", net.load.index)]
        super().__init__(net, [("gen", "p_mw", net.gen.index)], obs_keys, train_data="normal_around_mean",
                         test_data="normal_around_mean", **kwargs)


def test_non_simbench_example_steps_and_matches_the_oracle():
    env = NonSimbenchNet(num_envs=6, seed=2, obs_dtype="float64", engine_cls=TorchHostSimEngine)
    obs, _ = env.reset(seed=5)
    p0 = env.net.load.p_mw.to_numpy(float)
    got = obs[:, :len(p0)].numpy()
    assert (got >= p0 - 0.3 * np.abs(p0) - 1e-12).all() and (got <= p0 + 0.3 * np.abs(p0) + 1e-12).all()
    assert got.std(axis=0).min() > 0                                  # every load is sampled per environment
    act = torch.rand(6, len(env.net.gen), dtype=torch.float64, generator=torch.Generator().manual_seed(3))
    state = env.engine.state.clone()
    obs, reward, term, trunc, info = env.step(act)
    assert bool(info["converged"].all()) and term.all()
    lay = env.program.layout
    for b in range(6):
        net = env.net.deepcopy()
        net.load["p_mw"] = state[b, lay.slice("load", "p_mw")].numpy()
        net.load["q_mvar"] = state[b, lay.slice("load", "q_mvar")].numpy()
        net.gen["p_mw"] = act[b].numpy() * 80.0
        pf.runpp(net, enforce_q_lims=True)
        np.testing.assert_allclose(env.engine.vm[b].numpy()[env.program.ppc.bus_lookup], net.res_bus.vm_pu.to_numpy(),
                                   atol=1e-9)
        assert float(env.engine.objective[b]) == pytest.approx(-float(net.res_ext_grid.p_mw.iloc[0]), rel=1e-9)
