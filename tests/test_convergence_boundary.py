"""Converged flag and iteration count AT THE CONVERGENCE BOUNDARY (north_star: "the converged /
not-converged flag is bit-exact").

Every other parity batch converges to 100 %; this one does not.  The bus injections that kernel 1
produced are multiplied per environment by a factor drawn across the voltage-collapse boundary, so
that a large share of the batch diverges, runs into ``max_iteration`` = 10, or needs 6-9 iterations.
The engine (static-pivot 2x2-block LU on a fixed schedule) and the oracle (PYPOWER ``newtonpf`` on
SciPy's SuperLU with partial pivoting) then solve the SAME injections; flag and iteration count of
EVERY environment are compared, voltages of the converged ones.  Also in the batch: NaN injections
(a NaN action is data, opf_env.py:382 is not checked on the device) and -- on the meshed grid --
generators whose reactive limits bind, so that the PV->PQ outer loop runs next to divergence.

The reference's control flow around the flag: opfgym/opf_env.py:656-662 (exception -> False) and
:390-399 (what ``step`` emits for a failed power flow).
"""
import multiprocessing as mp
import os

import numpy as np
import pytest

from opfgym_b200 import ppc as P
from tests import common

_G = {}


def _oracle_one(b):
    """Environment b on the ORACLE's own tables (oracle/ppc_ref.py: its own per-unit conversion and bus
    numbering); only the injections come from the engine, carried over through the pandapower bus order."""
    import copy
    from oracle import pf
    ref, perm, sbus = _G["ref"], _G["perm"], _G["sbus"][b]
    if not np.isfinite(sbus).all():
        return False, -1, None, None
    s = np.zeros(ref.bus.shape[0], complex)
    s[perm] = sbus[:, 0] + 1j * sbus[:, 1]
    bus, gen = ref.bus.copy(), ref.gen
    on = gen[:, P.GEN_STATUS] > 0
    sg = np.zeros(len(bus), complex)
    np.add.at(sg, gen[on, P.GEN_BUS].astype(int), gen[on, P.PG] + 1j * gen[on, P.QG])
    sd = sg - s * ref.base_mva                     # makeSbus then returns exactly these injections
    bus[:, P.PD], bus[:, P.QD] = sd.real, sd.imag
    case = copy.copy(ref)
    case.bus = bus
    res = pf.run_pf(case, tolerance_mva=1e-8, max_iteration=10, enforce_q_lims=True, init="dc")
    types_changed = int((res["bus"][:, P.BUS_TYPE] != ref.bus[:, P.BUS_TYPE]).sum())
    return bool(res["converged"]), int(res["iterations"]), np.abs(res["V"])[perm], types_changed


def _oracle_batch(net, ppc, sbus, envs):
    from oracle import ppc_ref
    ref = ppc_ref.build(net)
    ok = ppc.bus_lookup >= 0
    perm = np.full(ppc.bus.shape[0], -1)           # engine bus -> oracle bus
    perm[ppc.bus_lookup[ok]] = ref.bus_lookup[ok]
    assert (perm >= 0).all(), "stand-in grids of this test have no auxiliary buses"
    _G["ref"], _G["perm"], _G["sbus"] = ref, perm, sbus
    n = min(len(os.sched_getaffinity(0)), 32)
    if n > 1 and len(envs) > 64:
        with mp.get_context("fork").Pool(n) as pool:
            return pool.map(_oracle_one, list(envs), chunksize=16)
    return [_oracle_one(b) for b in envs]


def _gen_limits(net):
    net.gen["min_q_mvar"] = -8.0
    net.gen["max_q_mvar"] = 8.0
    net.gen["vm_pu"] = np.linspace(0.99, 1.03, len(net.gen))


def _stress(engine_cls, name, n_env, lam, seed, sync=lambda: None, every=1):
    case = common.make_case(name, prepare=_gen_limits if "HV" in name else None)
    eng = engine_cls(case.program, n_env, obs_dtype="float64")
    common.randomize(case, eng, seed=seed)
    eng.assemble()
    sync()
    rng = np.random.default_rng(seed + 1)
    factor = rng.uniform(lam[0], lam[1], n_env)
    sbus = common._np(eng.sbus) * factor[:, None, None]
    nan_envs = rng.choice(n_env, size=max(2, n_env // 128), replace=False)
    sbus[nan_envs, rng.integers(0, sbus.shape[1], len(nan_envs)), 0] = np.nan
    eng.sbus[:] = sbus if isinstance(eng.sbus, np.ndarray) else eng._from_numpy(sbus)
    eng.pf_solve()
    sync()
    conv = common._np(eng.converged).astype(bool)
    iters = common._np(eng.iterations)
    vm = common._np(eng.vm)
    envs = range(0, n_env, every)
    ref = _oracle_batch(case.net, case.program.ppc, sbus, envs)
    flag_mismatch, iter_mismatch, worst_vm, switched = [], [], 0.0, 0
    for b, (ok, it, vm_ref, types_changed) in zip(envs, ref):
        if ok != conv[b]:
            flag_mismatch.append((b, factor[b], ok, it, int(iters[b])))
            continue
        if it >= 0 and it != iters[b]:
            iter_mismatch.append((b, factor[b], ok, it, int(iters[b])))
        if ok:
            worst_vm = max(worst_vm, float(np.abs(vm_ref - vm[b]).max()))
            switched += types_changed > 0
    share = conv[list(envs)].mean()
    return dict(flag_mismatch=flag_mismatch, iter_mismatch=iter_mismatch, worst_vm=worst_vm,
                converged_share=share, iters=np.bincount(iters[list(envs)], minlength=11),
                nan_flags=conv[nan_envs], switched=switched)


def _assert(out, need_switching=False):
    assert not out["flag_mismatch"], out["flag_mismatch"][:8]
    assert not out["iter_mismatch"], out["iter_mismatch"][:8]
    assert out["worst_vm"] < 1e-6, out["worst_vm"]
    assert 0.15 < out["converged_share"] < 0.85, out["converged_share"]      # really a mixed batch
    assert out["iters"][10] > 0 and out["iters"][5:10].sum() > 0, out["iters"]   # max_iter hits and slow convergers
    assert not out["nan_flags"].any()
    if need_switching:
        assert out["switched"] > 0


def test_boundary_hostsim_mv():
    from tests.hostsim.harness import HostSimEngine
    _assert(_stress(HostSimEngine, "1-MV-semiurb--1-sw", 192, (5.0, 12.0), seed=21))


def test_boundary_hostsim_hv_with_q_limits():
    from tests.hostsim.harness import HostSimEngine
    _assert(_stress(HostSimEngine, "1-HV-urban--0-sw", 64, (3.0, 9.0), seed=22), need_switching=True)


@pytest.mark.gpu
@pytest.mark.parametrize("name,n_env,lam,every,switching", [
    ("1-MV-semiurb--1-sw", 4096, (5.0, 12.0), 1, False),
    ("1-HV-urban--0-sw", 4096, (3.0, 9.0), 4, True),
])
def test_boundary_cuda(cuda_lib, name, n_env, lam, every, switching):
    import torch
    from opfgym_b200.engine import Engine
    out = _stress(Engine, name, n_env, lam, seed=23, sync=torch.cuda.synchronize, every=every)
    print(name, "converged share", out["converged_share"], "iterations", out["iters"].tolist(),
          "worst |dVm|", out["worst_vm"], "envs with PV->PQ switching", out["switched"])
    _assert(out, need_switching=switching)
