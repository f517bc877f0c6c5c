"""The oracle's own net -> ppc conversion (oracle/ppc_ref.py: element-wise loops, breadth-first bus
numbering) against the product's (opfgym_b200/ppc.py: vectorised, pandapower bus order) on every
stand-in grid: same buses, same injections, same branch parameters after mapping the numbering --
and, end to end, the same power-flow solution.  (Round-1 verdict: the oracle used to consume the
product's matrices, so line / transformer per-unit conversion, the wye-delta model, bus fusing and
the loading factors were checked against nothing.)"""
import numpy as np
import pytest

from opfgym_b200 import grids, ppc as P
from oracle import pf, ppc_ref as R

GRIDS = ["1-MV-semiurb--1-sw", "1-MV-rural--0-sw", "1-MV-comm--2-sw", "1-HV-urban--0-sw", "1-HV-mixed--1-sw"]


def _perturb(net, seed):
    rng = np.random.default_rng(seed)
    for t in ("load", "sgen", "storage"):
        if len(net[t]):
            net[t]["p_mw"] = net[t].p_mw.to_numpy() * rng.uniform(0.3, 1.2, len(net[t]))
    if len(net.trafo):
        net.trafo["tap_pos"] = rng.integers(-2, 3, len(net.trafo)).astype(float)


@pytest.mark.parametrize("name", GRIDS)
def test_builders_agree(name):
    net, _ = grids.build_simbench_net(name, n_profile_steps=96)
    _perturb(net, 1)
    a = P.PpcBuilder(net).build(net)
    b = R.build(net)
    assert a.bus.shape == b.bus.shape and a.branch.shape == b.branch.shape and a.gen.shape == b.gen.shape
    ok = a.bus_lookup >= 0
    assert np.array_equal(ok, b.bus_lookup >= 0)
    perm = np.full(a.bus.shape[0], -1)                      # product bus -> oracle bus
    perm[a.bus_lookup[ok]] = b.bus_lookup[ok]
    # auxiliary buses (half-open lines) have no pandapower bus: match them through their line
    for la, lb in zip(a.line_branch, b.line_branch):
        if la >= 0:
            for col in (P.F_BUS, P.T_BUS):
                perm[int(a.branch[la, col])] = int(b.branch[lb, col])
    assert (perm >= 0).all() and len(set(perm)) == len(perm)
    for col in (P.BUS_TYPE, P.PD, P.QD, P.GS, P.BS, P.VM, P.VA, P.BASE_KV):
        np.testing.assert_allclose(a.bus[:, col], b.bus[perm, col], rtol=1e-12, atol=1e-12, err_msg=f"bus col {col}")
    for ea, eb in ((a.line_branch, b.line_branch), (a.trafo_branch, b.trafo_branch)):
        assert np.array_equal(ea >= 0, eb >= 0)
        for ra, rb in zip(ea[ea >= 0], eb[eb >= 0]):
            assert perm[int(a.branch[ra, P.F_BUS])] == int(b.branch[rb, R.F_BUS])
            assert perm[int(a.branch[ra, P.T_BUS])] == int(b.branch[rb, R.T_BUS])
            for col in (P.BR_R, P.BR_X, P.BR_B, P.BR_G, P.TAP, P.SHIFT, P.BR_STATUS, P.RATE_A):
                assert a.branch[ra, col] == pytest.approx(b.branch[rb, col], rel=1e-12, abs=1e-15), col
            assert a.rate_f[ra] == pytest.approx(b.rate_f[rb], rel=1e-12)
            assert a.rate_t[ra] == pytest.approx(b.rate_t[rb], rel=1e-12)
    for ga, gb in ((a.ext_grid_gen, b.ext_grid_gen), (a.gen_gen, b.gen_gen)):
        assert np.array_equal(ga >= 0, gb >= 0)
        for ra, rb in zip(ga[ga >= 0], gb[gb >= 0]):
            assert perm[int(a.gen[ra, P.GEN_BUS])] == int(b.gen[rb, R.GEN_BUS])
            np.testing.assert_allclose(a.gen[ra, 1:], b.gen[rb, 1:], rtol=1e-12)


@pytest.mark.parametrize("name", GRIDS[:2] + GRIDS[3:4])
def test_same_solution_through_either_builder(name):
    net, _ = grids.build_simbench_net(name, n_profile_steps=96)
    _perturb(net, 2)
    one, two = net.deepcopy(), net.deepcopy()
    pf.runpp(one)                                  # oracle's builder
    pf.runpp(two, P.PpcBuilder(two))               # product's builder, same solver
    np.testing.assert_allclose(one.res_bus.vm_pu, two.res_bus.vm_pu, atol=1e-11)
    np.testing.assert_allclose(one.res_bus.va_degree, two.res_bus.va_degree, atol=1e-9)
    np.testing.assert_allclose(one.res_line.loading_percent, two.res_line.loading_percent, atol=1e-8)
    np.testing.assert_allclose(one.res_trafo.loading_percent, two.res_trafo.loading_percent, atol=1e-8)
    np.testing.assert_allclose(one.res_ext_grid.to_numpy(float), two.res_ext_grid.to_numpy(float), atol=1e-8)


def test_switches_fusing_and_islands():
    """Bus-bus switch fusion, a half-open line (auxiliary bus), a fully open line, a dead-end island."""
    net, _ = grids.build_simbench_net("1-MV-comm--2-sw", n_profile_steps=96)
    a = P.PpcBuilder(net).build(net)
    b = R.build(net)
    assert (a.bus_lookup >= 0).sum() == (b.bus_lookup >= 0).sum()
    assert a.bus.shape[0] == b.bus.shape[0]
    # buses fused by a closed bus-bus switch share one ppc bus in both builders
    sw = net.switch[(net.switch.et == "b") & net.switch.closed.astype(bool)]
    pos = {int(x): i for i, x in enumerate(net.bus.index)}
    for x, y in zip(sw.bus, sw.element):
        assert a.bus_lookup[pos[int(x)]] == a.bus_lookup[pos[int(y)]]
        assert b.bus_lookup[pos[int(x)]] == b.bus_lookup[pos[int(y)]]
