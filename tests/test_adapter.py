"""``power_flow_solver(net)`` plug-in contract (opf_env.py:53,70,646-662): fills net.res_*
like the oracle's runpp, raises LoadflowNotConverged otherwise.  CPU tier: host-sim."""
import numpy as np
import pytest

from opfgym_b200 import adapter, grids
from opfgym_b200.net import LoadflowNotConverged
from oracle import pf
from tests.hostsim.harness import TorchHostSimEngine


def _check(solver_kwargs):
    net, _ = grids.build_simbench_net("1-HV-urban--0-sw", n_profile_steps=96)
    solver = adapter.PowerFlowSolver(net, **solver_kwargs)
    ref = net.deepcopy()
    pf.runpp(ref)
    solver(net)
    assert net.converged
    np.testing.assert_allclose(net.res_bus.vm_pu, ref.res_bus.vm_pu, atol=1e-9)
    np.testing.assert_allclose(net.res_bus.va_degree, ref.res_bus.va_degree, atol=1e-7)
    for t in ("res_line", "res_trafo"):
        for c in ref[t].columns:
            np.testing.assert_allclose(net[t][c], ref[t][c], atol=1e-6, err_msg=f"{t}.{c}")
    np.testing.assert_allclose(net.res_ext_grid.to_numpy(), ref.res_ext_grid.to_numpy(), atol=1e-6)
    np.testing.assert_allclose(net.res_gen[["p_mw", "q_mvar"]].to_numpy(),
                               ref.res_gen[["p_mw", "q_mvar"]].to_numpy(), atol=1e-6)
    np.testing.assert_allclose(net.res_load.to_numpy(), ref.res_load.to_numpy(), atol=1e-12)
    # a second call with new injections reuses the compiled grid
    net.load["p_mw"] *= 0.5
    ref.load["p_mw"] *= 0.5
    solver(net)
    pf.runpp(ref)
    np.testing.assert_allclose(net.res_bus.vm_pu, ref.res_bus.vm_pu, atol=1e-9)
    net.load["p_mw"] *= 200.0
    with pytest.raises(LoadflowNotConverged):
        solver(net)
    assert not net.converged


def test_adapter_on_hostsim():
    _check(dict(engine_cls=TorchHostSimEngine))


@pytest.mark.gpu
def test_adapter_on_cuda(cuda_lib):
    _check({})


def test_unsupported_element_tables_fail_loudly():
    """A pandapower net with element types the builder does not model must not be solved silently."""
    import pandas as pd
    from opfgym_b200 import grids
    from opfgym_b200.ppc import PpcBuilder
    net, _ = grids.build_simbench_net("1-MV-semiurb--1-sw", n_profile_steps=96)
    PpcBuilder(net)                                            # fine as it is
    net.trafo3w = pd.DataFrame({"hv_bus": [0], "mv_bus": [1], "lv_bus": [2]})
    with pytest.raises(NotImplementedError, match="trafo3w"):
        PpcBuilder(net)
