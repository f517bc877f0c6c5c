"""``from_pandapower`` without pandapower: ``net._ppc`` / ``net._pd2ppc_lookups`` are fabricated in the layout the
adapter's notes describe for pandapower 2.13 / 2.14 ([ext-mem]: external-numbered ppc with dead buses of type 4 and
dead branches of status 0, one branch row per line / trafo / impedance in table order, lookup arrays) from the
ORACLE's own net -> ppc conversion, whose bus numbering (breadth-first) differs from the product's.  The adapter must
read that back into the tables the engine compiles -- checked against the product's own conversion after mapping the
numbering, and end to end: the single-net plug-in solves the same voltages through either builder.

This executes the reading logic here; that the layout IS pandapower's is what tests/parity/test_vs_pandapower.py
checks wherever pandapower is importable."""
import numpy as np
import pytest

from opfgym_b200 import adapter, grids, net as N, ppc as P
from opfgym_b200.pandapower_adapter import from_pandapower
from oracle import ppc_ref as R
from tests.hostsim.harness import TorchHostSimEngine as HostSimEngine


def _fabricate(net, complex_b):
    r = R.build(net)
    nb = r.bus.shape[0]
    dead = nb                                              # one more bus row: type 4, where dropped buses point
    bus = np.vstack([r.bus, np.zeros((1, r.bus.shape[1]))])
    bus[dead, R.BUS_I], bus[dead, R.BUS_TYPE], bus[dead, R.VM], bus[dead, R.BASE_KV] = dead, 4, 1.0, 1.0
    lookup = np.full(int(net.bus.index.max()) + 1, -1, dtype=np.int64)
    for idx, k in zip(net.bus.index, r.bus_lookup):
        lookup[int(idx)] = k if k >= 0 else dead
    rows, ranges = [], {}
    for table, mapping in (("line", r.line_branch), ("trafo", r.trafo_branch), ("impedance", r.impedance_branch)):
        first = len(rows)
        for k in mapping:
            if k >= 0:
                rows.append(r.branch[k].copy())
            else:                                          # out of service / cut off: a row stays, status 0
                row = np.zeros(r.branch.shape[1])
                row[R.F_BUS] = row[R.T_BUS] = dead
                row[R.TAP] = 1.0
                rows.append(row)
        if len(mapping):
            ranges[table] = (first, len(rows))
    branch = np.array(rows)
    if complex_b:                                          # pandapower < 2.14: y_shunt = 1j * BR_B, no BR_G column
        z = branch.astype(complex)
        z[:, R.BR_B] = branch[:, R.BR_B] - 1j * branch[:, R.BR_G]
        branch = z[:, :R.BR_G]
    net._ppc = {"baseMVA": r.base_mva, "bus": bus, "gen": r.gen.copy(), "branch": branch}
    net._pd2ppc_lookups = {"bus": lookup, "branch": ranges}
    return r


@pytest.fixture(autouse=True)
def _pandapower_column_index(monkeypatch):
    """The one constant the adapter takes from pandapower itself: where >= 2.14 keeps the branch conductance."""
    import sys, types
    pp, pyp, idx = types.ModuleType("pandapower"), types.ModuleType("pandapower.pypower"), types.ModuleType("pandapower.pypower.idx_brch")
    idx.BR_G = R.BR_G                                      # the fabricated table keeps it where the oracle does
    pp.pypower, pyp.idx_brch = pyp, idx
    for name, mod in (("pandapower", pp), ("pandapower.pypower", pyp), ("pandapower.pypower.idx_brch", idx)):
        monkeypatch.setitem(sys.modules, name, mod)


def _net():
    net, _ = grids.build_simbench_net("1-MV-comm--2-sw", n_profile_steps=96)      # fused buses, open ties, a dead end
    live = net.bus.index.to_numpy()[P.PpcBuilder(net).build(net).bus_lookup >= 0]
    vn = net.bus.vn_kv
    same = [b for b in live if vn[b] == vn[live[5]] and b != live[5]]
    N.create_impedance(net, live[5], same[-1], rft_pu=0.03, xft_pu=0.08, sn_mva=25.0)
    N.create_ward(net, live[7], ps_mw=0.2, qs_mvar=0.05, pz_mw=0.1, qz_mvar=-0.1)
    net.line.loc[net.line.index[3], "in_service"] = False
    return net


@pytest.mark.parametrize("complex_b", [False, True])
def test_adapter_reads_a_pandapower_shaped_ppc(complex_b):
    net = _net()
    _fabricate(net, complex_b)
    got = from_pandapower(net).build(net)
    own = P.PpcBuilder(net).build(net)
    assert got.bus.shape == own.bus.shape and got.branch.shape == own.branch.shape and got.gen.shape == own.gen.shape
    ok = own.bus_lookup >= 0
    assert np.array_equal(ok, got.bus_lookup >= 0)
    perm = np.full(own.bus.shape[0], -1)                   # product bus -> adapter bus
    perm[own.bus_lookup[ok]] = got.bus_lookup[ok]
    for la, lb in zip(own.line_branch, got.line_branch):   # auxiliary buses of half-open lines: through their line
        if la >= 0:
            for col in (P.F_BUS, P.T_BUS):
                perm[int(own.branch[la, col])] = int(got.branch[lb, col])
    assert (perm >= 0).all() and len(set(perm)) == len(perm)
    for col in (P.BUS_TYPE, P.GS, P.BS, P.BASE_KV, P.VM):
        np.testing.assert_allclose(own.bus[:, col], got.bus[perm, col], rtol=1e-12, atol=1e-12)
    for mine, theirs in ((own.line_branch, got.line_branch), (own.trafo_branch, got.trafo_branch)):
        assert np.array_equal(mine >= 0, theirs >= 0)
        for ra, rb in zip(mine[mine >= 0], theirs[theirs >= 0]):
            assert perm[int(own.branch[ra, P.F_BUS])] == int(got.branch[rb, P.F_BUS])
            for col in (P.BR_R, P.BR_X, P.BR_B, P.BR_G, P.TAP, P.SHIFT):
                assert own.branch[ra, col] == pytest.approx(got.branch[rb, col], rel=1e-12, abs=1e-15)
            assert own.rate_f[ra] == pytest.approx(got.rate_f[rb], rel=1e-12)
            assert own.rate_t[ra] == pytest.approx(got.rate_t[rb], rel=1e-12)
    # the impedance element came through as a branch without result rows
    builder = from_pandapower(net)
    assert list(builder.other_branch_rows) == ["impedance"] and len(builder.other_branch_rows["impedance"]) == 1
    k = int(builder.other_branch_rows["impedance"][0])
    assert got.branch[k, P.BR_R] == pytest.approx(0.03 * net.sn_mva / 25.0) and got.rate_f[k] == 0.0


def test_single_net_plug_in_through_the_adapter():
    net = _net()
    _fabricate(net, complex_b=False)
    a, b = net.deepcopy(), net.deepcopy()
    b._ppc, b._pd2ppc_lookups = net._ppc, net._pd2ppc_lookups
    adapter.PowerFlowSolver(a, engine_cls=HostSimEngine)(a)                       # product's own conversion
    adapter.PowerFlowSolver(b, engine_cls=HostSimEngine, builder=from_pandapower(b))(b)
    np.testing.assert_allclose(a.res_bus.vm_pu, b.res_bus.vm_pu, atol=1e-11, equal_nan=True)
    np.testing.assert_allclose(a.res_bus.va_degree, b.res_bus.va_degree, atol=1e-9, equal_nan=True)
    np.testing.assert_allclose(a.res_line.loading_percent, b.res_line.loading_percent, atol=1e-8, equal_nan=True)
    assert np.isfinite(a.res_bus.vm_pu.to_numpy()).sum() > 50


def test_xward_is_rejected():
    net = _net()
    _fabricate(net, complex_b=False)
    last = max(b for _, b in net._pd2ppc_lookups["branch"].values())
    net._ppc["branch"] = np.vstack([net._ppc["branch"], net._ppc["branch"][:1]])
    net._pd2ppc_lookups["branch"]["xward"] = (last, last + 1)
    with pytest.raises(NotImplementedError, match="xward"):
        from_pandapower(net)
