"""IEEE 14-bus test system: a second published known answer for the power flow, with what the
WSCC-9 case lacks -- off-nominal tap transformers, a bus shunt, five voltage-controlled buses.

Case data and solution are the ones published with MATPOWER (`case14`, `runpf`; values to the
printed precision).  The oracle (oracle/pf.py), the host build of the C ABI and the CUDA library are
all held against them; the library is driven through `opfg_grid_create` / `opfg_pf_solve` only, the
way INTEGRATION.md §3 shows a foreign host doing it.
"""
import ctypes as C

import numpy as np
import pytest

from opfgym_b200 import capi
from opfgym_b200 import ppc as P
from oracle import pf

LOADS = {2: (21.7, 12.7), 3: (94.2, 19.0), 4: (47.8, -3.9), 5: (7.6, 1.6), 6: (11.2, 7.5), 9: (29.5, 16.6),
         10: (9.0, 5.8), 11: (3.5, 1.8), 12: (6.1, 1.6), 13: (13.5, 5.8), 14: (14.9, 5.0)}
GENS = [(1, 232.4, 1.06, 10, 0), (2, 40, 1.045, 50, -40), (3, 0, 1.01, 40, 0), (6, 0, 1.07, 24, -6),
        (8, 0, 1.09, 24, -6)]                                   # bus, Pg, Vg, Qmax, Qmin
BRANCHES = [(1, 2, .01938, .05917, .0528, 0), (1, 5, .05403, .22304, .0492, 0), (2, 3, .04699, .19797, .0438, 0),
            (2, 4, .05811, .17632, .0340, 0), (2, 5, .05695, .17388, .0346, 0), (3, 4, .06701, .17103, .0128, 0),
            (4, 5, .01335, .04211, 0, 0), (4, 7, 0, .20912, 0, .978), (4, 9, 0, .55618, 0, .969),
            (5, 6, 0, .25202, 0, .932), (6, 11, .09498, .19890, 0, 0), (6, 12, .12291, .25581, 0, 0),
            (6, 13, .06615, .13027, 0, 0), (7, 8, 0, .17615, 0, 0), (7, 9, 0, .11001, 0, 0),
            (9, 10, .03181, .08450, 0, 0), (9, 14, .12711, .27038, 0, 0), (10, 11, .08205, .19207, 0, 0),
            (12, 13, .22092, .19988, 0, 0), (13, 14, .17093, .34802, 0, 0)]   # f, t, r, x, b, tap ratio
# published solution
VM14 = [1.060, 1.045, 1.010, 1.0177, 1.0195, 1.070, 1.0615, 1.090, 1.0559, 1.0510, 1.0569, 1.0552, 1.0504, 1.0355]
VA14 = [0.0, -4.983, -12.725, -10.313, -8.774, -14.221, -13.360, -13.360, -14.939, -15.097, -14.791, -15.076,
        -15.156, -16.034]
PG_SLACK, QG14 = 232.39, [-16.55, 43.56, 25.08, 12.73, 17.62]


def case14():
    nb = 14
    bus = np.zeros((nb, P.BUS_COLS))
    bus[:, P.BUS_I] = np.arange(nb)
    bus[:, P.BUS_TYPE] = P.PQ
    bus[:, P.VM] = 1.0
    bus[:, P.BASE_KV] = 135.0
    for b, (p, q) in LOADS.items():
        bus[b - 1, [P.PD, P.QD]] = [p, q]
    bus[8, P.BS] = 19.0
    gen = np.zeros((len(GENS), P.GEN_COLS))
    for i, (b, pg, vg, qmax, qmin) in enumerate(GENS):
        gen[i, [P.GEN_BUS, P.PG, P.VG, P.QMAX, P.QMIN, P.GEN_STATUS]] = [b - 1, pg, vg, qmax, qmin, 1]
        bus[b - 1, P.VM] = vg
        bus[b - 1, P.BUS_TYPE] = P.REF if i == 0 else P.PV
    branch = np.zeros((len(BRANCHES), P.BRANCH_COLS))
    for i, (f, t, r, x, b, tap) in enumerate(BRANCHES):
        branch[i, :5] = [f - 1, t - 1, r, x, b]
        branch[i, P.TAP] = tap
    branch[:, P.BR_STATUS] = 1
    return P.Ppc(100.0, bus, gen, branch, np.arange(nb), np.arange(nb), np.zeros(0, int), np.array([0]),
                 np.array([1, 2, 5, 7]), np.ones(nb), np.ones(len(BRANCHES)))


def check_solution(vm, va_deg):
    np.testing.assert_allclose(vm, VM14, rtol=0, atol=6e-5)        # published to 4 decimals
    np.testing.assert_allclose(va_deg, VA14, rtol=0, atol=6e-4)    # published to 3 decimals


@pytest.mark.parametrize("init", ["flat", "dc"])
def test_oracle_matches_published_ieee14(init):
    res = pf.run_pf(case14(), tolerance_mva=1e-6, init=init, enforce_q_lims=False)
    assert res["converged"] and res["iterations"] <= 4
    check_solution(np.abs(res["V"]), np.degrees(np.angle(res["V"])))
    g = res["gen"]
    assert abs(g[0, P.PG] - PG_SLACK) < 6e-3
    np.testing.assert_allclose(g[:, P.QG], QG14, rtol=0, atol=6e-3)
    assert abs(g[:, P.PG].sum() - 259.0 - 13.393) < 2e-3           # published losses 13.393 MW


def solve_through_c_abi(lib, make_buffer, to_numpy, sync=lambda: None, n_env=3):
    """opfg_grid_create + opfg_pf_solve on `n_env` copies of the case (Sbus = makeSbus of the ppc)."""
    ppc = case14()
    bus, gen, branch = (np.ascontiguousarray(a, dtype=np.float64) for a in (ppc.bus, ppc.gen, ppc.branch))
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    desc = capi.GridDesc(nb=14, ng=len(GENS), nbr=len(BRANCHES), base_mva=100.0, bus=dp(bus), bus_cols=bus.shape[1],
                         gen=dp(gen), gen_cols=gen.shape[1], branch=dp(branch), branch_cols=branch.shape[1],
                         tol_pu=1e-8, max_iter=10, init_dc=1, enforce_q_lims=0, threads_per_env=0, ordering=0)
    handle = C.c_void_p()
    capi.check(lib, lib.opfg_grid_create(C.byref(desc), C.byref(handle)))
    try:
        sbus = pf.make_sbus(ppc.base_mva, ppc.bus, ppc.gen)          # complex [nb], per unit
        host = np.tile(np.stack([sbus.real, sbus.imag], axis=1)[None], (n_env, 1, 1))
        d_sbus = make_buffer(host)
        d_vm, d_va = make_buffer(np.zeros((n_env, 14))), make_buffer(np.zeros((n_env, 14)))
        d_conv = make_buffer(np.zeros(n_env, np.uint8))
        d_it = make_buffer(np.zeros(n_env, np.int32))
        ptr = lambda t: C.c_void_p(t.data_ptr() if hasattr(t, "data_ptr") else t.ctypes.data)
        batch = capi.Batch(n_env=n_env, sbus=ptr(d_sbus), vm=ptr(d_vm), va=ptr(d_va), converged=ptr(d_conv),
                           iterations=ptr(d_it))
        capi.check(lib, lib.opfg_pf_solve(handle, C.byref(batch), None))
        sync()
        return to_numpy(d_vm), to_numpy(d_va), to_numpy(d_conv), to_numpy(d_it)
    finally:
        lib.opfg_grid_destroy(handle)


def test_host_build_of_c_abi_matches_published_ieee14():
    from tests.hostsim import harness
    vm, va, conv, it = solve_through_c_abi(harness.load(), lambda a: np.ascontiguousarray(a).copy(), np.asarray)
    assert conv.all() and (it <= 4).all()
    for b in range(vm.shape[0]):
        check_solution(vm[b], np.degrees(va[b]))


@pytest.mark.gpu
def test_cuda_library_matches_published_ieee14(cuda_lib):
    import torch
    vm, va, conv, it = solve_through_c_abi(
        cuda_lib, lambda a: torch.as_tensor(np.ascontiguousarray(a)).cuda(), lambda t: t.cpu().numpy(),
        sync=torch.cuda.synchronize, n_env=300)
    assert conv.all() and (it <= 4).all()
    for b in (0, 150, 299):
        check_solution(vm[b], np.degrees(va[b]))
    assert np.ptp(vm, axis=0).max() == 0.0 and np.ptp(va, axis=0).max() == 0.0   # identical for every copy
