"""The two power-flow kernels -- one CTA per environment (level-scheduled block LU in shared memory)
and one LANE per environment (row-wise LU, Jacobian never stored) -- walk the same factorisation in the
same order of operations: every output bit must agree, on converged, slow and diverging environments."""
import numpy as np
import pytest

from tests import common


def _pair(engine_cls, name, n_env, seed, lam=None, sync=lambda: None, **kw):
    case = common.make_case(name)
    out = {}
    for kernel in ("cta", "lanes"):
        # same elimination order for both (3 = least fill work, what the lane kernel picks by itself)
        eng = engine_cls(case.program, n_env, obs_dtype="float64", pf_kernel=kernel, ordering=3, **kw)
        assert eng.info["pf_lanes"] == (kernel == "lanes"), eng.info
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
    a, b = out["cta"], out["lanes"]
    if mixed:
        assert 0.1 < a["converged"].mean() < 0.9
    else:
        assert a["converged"].all()
    for k in ("converged", "iterations"):
        assert np.array_equal(a[k], b[k]), k
    ok = a["converged"].astype(bool)
    for k in ("vm", "va", "reward", "obs"):
        assert np.array_equal(a[k][ok].view(np.int64), b[k][ok].view(np.int64)), k
    # diverging environments: the iterates agree bit for bit as well (NaN patterns included)
    assert np.array_equal(a["vm"].view(np.int64), b["vm"].view(np.int64))


@pytest.mark.parametrize("name", ["1-MV-semiurb--1-sw", "1-HV-urban--0-sw"])
def test_lanes_equal_cta_hostsim(name):
    from tests.hostsim.harness import HostSimEngine
    _same(_pair(HostSimEngine, name, 24, seed=31), mixed=False)


def test_lanes_equal_cta_at_the_convergence_boundary_hostsim():
    from tests.hostsim.harness import HostSimEngine
    _same(_pair(HostSimEngine, "1-MV-semiurb--1-sw", 96, seed=32, lam=(5.0, 12.0)), mixed=True)


@pytest.mark.gpu
@pytest.mark.parametrize("name,n_env,lam", [("1-MV-semiurb--1-sw", 4133, None),
                                            ("1-MV-semiurb--1-sw", 4096, (5.0, 12.0)),
                                            ("1-HV-urban--0-sw", 1000, None),
                                            ("1-HV-urban--0-sw", 1024, (3.0, 9.0))])
def test_lanes_equal_cta_cuda(cuda_lib, name, n_env, lam):
    import torch
    from opfgym_b200.engine import Engine
    _same(_pair(Engine, name, n_env, seed=33, lam=lam, sync=torch.cuda.synchronize), mixed=lam is not None)
