"""Pins the CPU oracle's power flow (oracle/pf.py) on known answers: the WSCC
9-bus textbook case, the 2-bus closed form (C.5) and the
invariants of C.6.  The pandapower boundary itself is unpinned (no PF number in
the reference's tests)."""
import numpy as np
import pytest

from opfgym_b200 import grids
from opfgym_b200 import ppc as P
from oracle import pf


def wscc9():
    bus = np.zeros((9, P.BUS_COLS))
    bus[:, P.BUS_I] = np.arange(9)
    bus[:, P.BUS_TYPE] = P.PQ
    bus[:, P.VM] = 1.0
    bus[:, P.BASE_KV] = 345.0
    bus[0, P.BUS_TYPE] = P.REF
    bus[[1, 2], P.BUS_TYPE] = P.PV
    bus[4, [P.PD, P.QD]] = [90, 30]
    bus[6, [P.PD, P.QD]] = [100, 35]
    bus[8, [P.PD, P.QD]] = [125, 50]
    gen = np.zeros((3, P.GEN_COLS))
    gen[:, P.GEN_BUS] = [0, 1, 2]
    gen[:, P.PG] = [0, 163, 85]
    gen[:, P.VG] = [1.04, 1.025, 1.025]
    gen[:, P.GEN_STATUS] = 1
    gen[:, P.QMAX], gen[:, P.QMIN] = 300, -300
    rows = [(1, 4, 0, .0576, 0), (4, 5, .017, .092, .158), (5, 6, .039, .17, .358),
            (3, 6, 0, .0586, 0), (6, 7, .0119, .1008, .209), (7, 8, .0085, .072, .149),
            (8, 2, 0, .0625, 0), (8, 9, .032, .161, .306), (9, 4, .01, .085, .176)]
    branch = np.zeros((9, P.BRANCH_COLS))
    for i, (f, t, r, x, b) in enumerate(rows):
        branch[i, :5] = [f - 1, t - 1, r, x, b]
    branch[:, P.BR_STATUS] = 1
    bus[[0, 1, 2], P.VM] = [1.04, 1.025, 1.025]
    return P.Ppc(100.0, bus, gen, branch, np.arange(9), np.arange(9), np.zeros(0, int),
                 np.array([0]), np.array([1, 2]), np.ones(9), np.ones(9))


VM9 = [1.04, 1.025, 1.025, 1.0257883928, 1.012654324, 1.032352949, 1.0158825836,
       1.0257693724, 0.995630858]
VA9 = [0, 9.2800054816, 4.6647513331, -2.2167877999, -3.6873961702, 1.9667160744,
       0.7275360769, 3.7197011546, -3.9888052729]


def test_wscc9_known_answer():
    res = pf.run_pf(wscc9(), tolerance_mva=1e-6, init="flat", enforce_q_lims=False)
    assert res["converged"] and res["iterations"] == 4
    np.testing.assert_allclose(np.abs(res["V"]), VM9, atol=5e-10)
    np.testing.assert_allclose(np.degrees(np.angle(res["V"])), VA9, atol=5e-9)
    g = res["gen"]
    np.testing.assert_allclose(g[0, [P.PG, P.QG]], [71.6410214745, 27.0459235335], atol=1e-8)
    np.testing.assert_allclose(g[1:, P.QG], [6.6536603184, -10.859709071], atol=1e-8)
    losses = g[:, P.PG].sum() - 315.0
    assert abs(losses - 4.6410214745) < 1e-8


def test_two_bus_closed_form():
    r, x, p, q = 0.02, 0.06, 0.8, 0.3
    bus = np.zeros((2, P.BUS_COLS))
    bus[:, P.BUS_TYPE] = [P.REF, P.PQ]
    bus[:, P.VM] = 1.0
    bus[1, [P.PD, P.QD]] = [p, q]
    gen = np.zeros((1, P.GEN_COLS))
    gen[0, [P.VG, P.GEN_STATUS]] = [1.0, 1]
    branch = np.zeros((1, P.BRANCH_COLS))
    branch[0, :4] = [0, 1, r, x]
    branch[0, P.BR_STATUS] = 1
    ppc = P.Ppc(1.0, bus, gen, branch, np.arange(2), np.arange(1), np.zeros(0, int),
                np.array([0]), np.zeros(0, int), np.ones(1), np.ones(1))
    res = pf.run_pf(ppc, init="flat")
    # V2^4 + (2(rP+xQ) - 1) V2^2 + (r^2+x^2)(P^2+Q^2) = 0
    b = 2 * (r * p + x * q) - 1
    c = (r * r + x * x) * (p * p + q * q)
    v2 = np.sqrt((-b + np.sqrt(b * b - 4 * c)) / 2)
    assert res["converged"] and res["iterations"] == 3
    assert abs(abs(res["V"][1]) - v2) < 1e-10
    assert abs(v2 - 0.9637719383784054) < 1e-12
    assert abs(np.angle(res["V"][1]) - (-0.0435925798)) < 1e-9


@pytest.mark.parametrize("name", ["1-MV-rural--0-sw", "1-HV-mixed--1-sw"])
def test_invariants_on_standin_grid(name):
    net, _ = grids.build_simbench_net(name, n_profile_steps=96)
    res = pf.runpp(net)
    ppc = res["ppc"]
    v = res["V"]
    sbus = pf.make_sbus(ppc.base_mva, res["bus"], res["gen"])
    mis = v * np.conj(res["Ybus"] @ v) - sbus
    nonref = ppc.bus[:, P.BUS_TYPE] != P.REF
    assert np.abs(mis[nonref]).max() < 1e-8
    # power balance: injections = branch losses + shunt consumption
    s_inj = (v * np.conj(res["Ybus"] @ v)).sum() * ppc.base_mva
    losses = (res["Sf"] + res["St"]).sum()
    shunt = (np.abs(v) ** 2 * np.conj(ppc.bus[:, P.GS] + 1j * ppc.bus[:, P.BS])).sum()
    assert abs(s_inj - losses - shunt) < 1e-7
    assert ((res["Sf"] + res["St"]).real > -1e-9).all()      # passive branches do not generate P
    assert np.isfinite(net.res_line.loading_percent[net.line.in_service]).all()


def test_nonconvergence_is_reported():
    net, _ = grids.build_simbench_net("1-MV-rural--0-sw", n_profile_steps=96)
    net.load["p_mw"] *= 60.0          # far beyond the voltage-collapse point
    from opfgym_b200.net import LoadflowNotConverged
    with pytest.raises(LoadflowNotConverged):
        pf.runpp(net)
