"""`only_worst_case_violations` with BOTH bounds of one constraint violated: the reference adds the
worst upper-bound violation and the worst lower-bound violation (constraints.py:77-80, 113-122);
the kernel used to keep one running maximum over both (round-1 advisor finding)."""
import numpy as np
import pytest

from tests import common
from tests.hostsim.harness import HostSimEngine


def _case():
    # voltage band 0.99-1.01 p.u.: the feeder ends sag below it while the busbars sit above it
    return common.make_case("1-MV-rural--0-sw", tight=True,
                            constraint_kwargs=dict(only_worst_case_violations=True, penalty_power=2.0))


def _check(eng, case, envs):
    vm_c = 0                      # the voltage constraint is discovered first
    lo, hi = case.net.bus.min_vm_pu.to_numpy(), case.net.bus.max_vm_pu.to_numpy()
    lk = case.program.ppc.bus_lookup
    vm = common._np(eng.vm)[:, lk]
    two_sided = ((vm > hi).any(axis=1) & (vm < lo).any(axis=1))
    assert two_sided[list(envs)].any(), "fixture must violate both bounds in at least one env"
    worst = common.compare_with_oracle(case, eng, envs)
    assert worst["flag_mismatch"] == 0 and worst["valid_mismatch"] == 0, worst
    assert worst["violation"] <= 1e-9 and worst["reward_rel"] <= 1e-9, worst
    # the two-sided sum is strictly larger than the single maximum the old kernel returned
    b = int(np.nonzero(two_sided)[0][0])
    single_max = max((vm[b] - hi).max(), (lo - vm[b]).max())
    assert common._np(eng.violations)[b, vm_c] > single_max * (1 + 1e-9)


def test_two_sided_worst_case_hostsim():
    case = _case()
    eng = HostSimEngine(case.program, 16, obs_dtype="float64")
    common.randomize(case, eng, seed=11)
    eng.step()
    _check(eng, case, range(16))


@pytest.mark.gpu
def test_two_sided_worst_case_cuda(cuda_lib):
    import torch
    from opfgym_b200.engine import Engine
    case = _case()
    eng = Engine(case.program, 64, obs_dtype="float64")
    common.randomize(case, eng, seed=11)
    eng.step()
    torch.cuda.synchronize()
    _check(eng, case, range(0, 64, 4))
