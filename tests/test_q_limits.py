"""enforce_q_lims: PV -> PQ switching outer loop of
pandapower's `_run_ac_pf_with_qlims_enforced`, per environment, on the device."""
import numpy as np
import pytest

from opfgym_b200 import adapter, grids
from oracle import pf
from tests.hostsim.harness import TorchHostSimEngine


def _net(qlim):
    net, _ = grids.build_simbench_net("1-HV-urban--0-sw", n_profile_steps=96)
    net.gen["min_q_mvar"] = -qlim
    net.gen["max_q_mvar"] = qlim
    net.gen["vm_pu"] = np.linspace(0.99, 1.03, len(net.gen))
    return net


def _check(engine_kwargs):
    for qlim, expect_binding in ((1e4, False), (8.0, True)):
        net = _net(qlim)
        ref = net.deepcopy()
        res = pf.runpp(ref, enforce_q_lims=True)
        solver = adapter.PowerFlowSolver(net, **engine_kwargs)
        assert (solver.engine.info["nb"], solver.program.ppc.gen.shape[0]) == (372, 6)
        solver(net)
        q = ref.res_gen.q_mvar.to_numpy()
        binding = np.isclose(np.abs(q), qlim, atol=1e-6)
        assert binding.any() == expect_binding
        np.testing.assert_allclose(net.res_bus.vm_pu, ref.res_bus.vm_pu, atol=1e-9)
        np.testing.assert_allclose(net.res_bus.va_degree, ref.res_bus.va_degree, atol=1e-7)
        np.testing.assert_allclose(net.res_gen.q_mvar, q, atol=1e-6)
        np.testing.assert_allclose(net.res_ext_grid.to_numpy(), ref.res_ext_grid.to_numpy(), atol=1e-6)
        if expect_binding:   # the limited buses float: their |V| left the set-point
            free = ref.res_gen.vm_pu.to_numpy()[binding]
            assert (np.abs(free - net.gen.vm_pu.to_numpy()[binding]) > 1e-5).all()
            without = net.deepcopy()
            pf.runpp(without, enforce_q_lims=False)
            assert np.abs(without.res_bus.vm_pu - ref.res_bus.vm_pu).max() > 1e-5


def test_q_limits_hostsim():
    _check(dict(engine_cls=TorchHostSimEngine))


def test_both_zero_limits_are_skipped_like_pandapower():
    net = _net(0.0)      # EcoDispatch sets both limits to 0 (envs/eco_dispatch.py:86-88)
    ref = net.deepcopy()
    pf.runpp(ref, enforce_q_lims=True)
    solver = adapter.PowerFlowSolver(net, engine_cls=TorchHostSimEngine)
    solver(net)
    np.testing.assert_allclose(net.res_gen.vm_pu if "vm_pu" in net.res_gen else net.gen.vm_pu,
                               net.gen.vm_pu, atol=1e-12)
    np.testing.assert_allclose(net.res_bus.vm_pu, ref.res_bus.vm_pu, atol=1e-9)
    assert np.abs(ref.res_gen.q_mvar).max() > 1.0     # PV buses still regulate


@pytest.mark.gpu
def test_q_limits_cuda(cuda_lib):
    _check({})
