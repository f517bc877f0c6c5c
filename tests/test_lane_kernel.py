"""The two power-flow kernels -- one CTA per environment (level-scheduled block LU in shared memory)
and one LANE per environment (row-wise LU, Jacobian never stored) -- walk the same factorisation in the
same order of operations: every output bit must agree, on converged, slow and diverging environments."""
import numpy as np
import pytest

from tests import common


def _pair(engine_cls, name, n_env, seed, lam=None, sync=lambda: None, kernels=("cta", "lanes"), **kw):
    case = common.make_case(name)
    out = {}
    for kernel in kernels:
        # same elimination order for all (1 = minimum degree: leaf-first on radial grids)
        eng = engine_cls(case.program, n_env, obs_dtype="float64", pf_kernel=kernel, ordering=1, **kw)
        assert eng.info["pf_kernel_used"] == {"cta": 1, "lanes": 2, "radial": 3}[kernel], eng.info
        common.randomize(case, eng, seed=seed)
        eng.assemble()
        sync()
        if lam is not None:
            f = np.random.default_rng(seed).uniform(lam[0], lam[1], n_env)
            sb = common._np(eng.sbus) * f[:, None, None]
            eng.sbus[:] = sb if isinstance(eng.sbus, np.ndarray) else eng._from_numpy(sb)
        eng.pf_solve()
        eng.score()
        sync()
        out[kernel] = {k: common._np(getattr(eng, k)).copy() for k in
                       ("vm", "va", "converged", "iterations", "reward", "obs")}
        out[kernel]["launches"] = eng.launch_count()
    return out


def _same(out, mixed):
    names = list(out)
    for other in names[1:]:
        # the radial kernel forms sum(L W) before subtracting it from J_kk: last bits differ
        _same2(out[names[0]], out[other], mixed, exact=other != "radial")


def _same2(a, b, mixed, exact=True):
    if mixed:
        assert 0.1 < a["converged"].mean() < 0.9
    else:
        assert a["converged"].all()
    for k in ("converged", "iterations"):
        assert np.array_equal(a[k], b[k]), k
    ok = a["converged"].astype(bool)
    if not exact:
        for k, tol in (("vm", 1e-11), ("va", 1e-11), ("reward", 1e-9), ("obs", 1e-8)):
            np.testing.assert_allclose(a[k][ok], b[k][ok], rtol=0, atol=tol, err_msg=k)
        return
    for k in ("vm", "va", "reward", "obs"):
        assert np.array_equal(a[k][ok].view(np.int64), b[k][ok].view(np.int64)), k
    # diverging environments: the iterates agree bit for bit as well (NaN patterns included)
    assert np.array_equal(a["vm"].view(np.int64), b["vm"].view(np.int64))


ALL = ("cta", "lanes", "radial")


@pytest.mark.parametrize("name,kernels", [("1-MV-semiurb--1-sw", ALL), ("1-MV-rural--0-sw", ALL),
                                          ("1-HV-urban--0-sw", ("cta", "lanes"))])
def test_kernels_agree_hostsim(name, kernels):
    from tests.hostsim.harness import HostSimEngine
    _same(_pair(HostSimEngine, name, 24, seed=31, kernels=kernels), mixed=False)


def test_kernels_agree_at_the_convergence_boundary_hostsim():
    from tests.hostsim.harness import HostSimEngine
    _same(_pair(HostSimEngine, "1-MV-semiurb--1-sw", 96, seed=32, lam=(5.0, 12.0), kernels=ALL), mixed=True)


def test_auto_picks_the_radial_kernel_on_radial_grids_only():
    from tests.hostsim.harness import HostSimEngine
    for name, used in (("1-MV-semiurb--1-sw", 3), ("1-HV-urban--0-sw", 1)):
        case = common.make_case(name)
        eng = HostSimEngine(case.program, 2)
        assert eng.info["pf_kernel_used"] == used, (name, eng.info)


@pytest.mark.gpu
@pytest.mark.parametrize("name,n_env,lam,kernels", [("1-MV-semiurb--1-sw", 4133, None, ALL),
                                                    ("1-MV-semiurb--1-sw", 4096, (5.0, 12.0), ALL),
                                                    ("1-MV-rural--0-sw", 2000, None, ALL),
                                                    ("1-HV-urban--0-sw", 1000, None, ("cta", "lanes")),
                                                    ("1-HV-urban--0-sw", 1024, (5.0, 14.0), ("cta", "lanes"))])
def test_kernels_agree_cuda(cuda_lib, name, n_env, lam, kernels):
    import torch
    from opfgym_b200.engine import Engine
    _same(_pair(Engine, name, n_env, seed=33, lam=lam, sync=torch.cuda.synchronize, kernels=kernels),
          mixed=lam is not None)


@pytest.mark.gpu
def test_block_kernel_builds_by_block_size_cuda(cuda_lib, monkeypatch):
    """k_pf_multi exists in three builds (launch bounds 768 / 384 / 256 threads: 80 / ~142 / ~144 registers).  The MV
    grid normally runs the 768 build (11 environments x 64 threads); capped at 6 and at 4 environments per CTA it
    runs the 384 and the 256 build: same bits as the lane kernel in all three."""
    import torch
    from opfgym_b200.engine import Engine
    for cap in (None, "6", "4"):
        if cap is None:
            monkeypatch.delenv("OPFG_ENVS_PER_CTA", raising=False)
        else:
            monkeypatch.setenv("OPFG_ENVS_PER_CTA", cap)
        _same(_pair(Engine, "1-MV-semiurb--1-sw", 2051, seed=35, lam=(5.0, 12.0), sync=torch.cuda.synchronize,
                    kernels=("cta", "lanes")), mixed=True)


def _slots_on_off(engine_cls, n_env, monkeypatch, sync=lambda: None, **kw):
    """The meshed-grid kernel with fill blocks sharing the storage slots of dead blocks (default) and with one slot
    per block (OPFG_SHARE_SLOTS=0): only addresses differ, so every output bit must agree -- on converged, slow and
    diverging environments (injections scaled across the collapse point)."""
    case = common.make_case("1-HV-urban--0-sw")
    out, smem = {}, {}
    for share in ("1", "0"):
        monkeypatch.setenv("OPFG_SHARE_SLOTS", share)
        eng = engine_cls(case.program, n_env, obs_dtype="float64", pf_kernel="cta", **kw)
        assert eng.info["pf_kernel_used"] == 1 and eng.info["n_fill_blocks"] > 0
        smem[share] = eng.info["smem_bytes_pf"]
        common.randomize(case, eng, seed=41)
        eng.assemble()
        sync()
        f = np.random.default_rng(41).uniform(0.5, 40.0, n_env)
        sb = common._np(eng.sbus) * f[:, None, None]
        eng.sbus[:] = sb if isinstance(eng.sbus, np.ndarray) else eng._from_numpy(sb)
        eng.pf_solve()
        eng.score()
        sync()
        out[share] = {k: common._np(getattr(eng, k)).copy() for k in ("vm", "va", "converged", "iterations", "reward")}
    monkeypatch.delenv("OPFG_SHARE_SLOTS")
    assert smem["1"] < 0.82 * smem["0"], smem                       # 1 899 blocks -> 1 347 slots on this grid
    conv = out["1"]["converged"].astype(bool)
    assert 0 < conv.sum() < n_env                                   # both kinds present
    for k in out["1"]:
        a, b = out["1"][k], out["0"][k]
        assert np.array_equal(a, b, equal_nan=a.dtype.kind == "f"), k


def test_shared_block_slots_change_no_bit_hostsim(monkeypatch):
    from tests.hostsim.harness import HostSimEngine
    _slots_on_off(HostSimEngine, 24, monkeypatch)


@pytest.mark.gpu
def test_shared_block_slots_change_no_bit_cuda(cuda_lib, monkeypatch):
    import torch
    from opfgym_b200.engine import Engine
    _slots_on_off(Engine, 1024, monkeypatch, sync=torch.cuda.synchronize)
