"""Per-environment voltage set-points ("PV buses with per-env vm_pu setpoints";
reference anchor opfgym/envs/eco_dispatch.py:83): ``gen.vm_pu`` as an ACTION column -- every
environment holds its own PV-bus voltages; checked per environment against the oracle, with and
without reactive limits that bind."""
import numpy as np
import pytest
import torch

from opfgym_b200 import grids
from opfgym_b200.opf_env import BatchedOpfEnv
from oracle import pf
from tests.hostsim.harness import TorchHostSimEngine


def _env(n, qlim, **kw):
    net, profiles = grids.build_simbench_net("1-HV-urban--0-sw", n_profile_steps=96)
    net.gen["min_vm_pu"] = 0.97
    net.gen["max_vm_pu"] = 1.04
    net.gen["min_q_mvar"] = -qlim
    net.gen["max_q_mvar"] = qlim
    obs_keys = [("load", "p_mw", net.load.index), ("load", "q_mvar", net.load.index),
                ("res_bus", "vm_pu", net.bus.index[:8])]
    act_keys = [("gen", "vm_pu", net.gen.index)]
    return BatchedOpfEnv(net, act_keys, obs_keys, profiles=profiles, num_envs=n, train_data="full_uniform",
                         test_data="full_uniform", seed=2, obs_dtype="float64", **kw)


def _check(qlim, expect_binding, sync=lambda: None, **kw):
    n = 6
    env = _env(n, qlim, **kw)
    assert env.engine.program.assembly["bus_vm_ref"] is not None
    env.reset(seed=4)
    act = torch.rand(n, len(env.net.gen), dtype=torch.float64, generator=torch.Generator().manual_seed(3))
    act[0] = 0.0; act[1] = 1.0                                   # all at the lower / upper set-point
    e = env.engine
    e.actions.copy_(act.to(env.device))
    state = e.state.clone()
    e.step()
    sync()
    assert bool(e.converged.all())
    set_points = env.col("gen", "vm_pu").cpu().numpy()
    assert np.ptp(set_points, axis=0).min() > 0.01               # really per environment
    lk = env.program.ppc.bus_lookup
    bound = 0
    for b in range(n):
        net = env.net.deepcopy()
        for t, c in (("load", "p_mw"), ("load", "q_mvar"), ("sgen", "p_mw"), ("storage", "p_mw")):
            if env.program.layout.has(t, c):
                net[t][c] = state[b, env.program.layout.slice(t, c)].cpu().numpy()
        net.gen["vm_pu"] = 0.97 + act[b].numpy() * (1.04 - 0.97)
        np.testing.assert_allclose(set_points[b], net.gen.vm_pu.to_numpy(), atol=1e-15)
        pf.runpp(net, enforce_q_lims=True)
        np.testing.assert_allclose(e.vm[b].cpu().numpy()[lk], net.res_bus.vm_pu.to_numpy(), atol=1e-9)
        np.testing.assert_allclose(np.degrees(e.va[b].cpu().numpy()[lk]), net.res_bus.va_degree.to_numpy(), atol=1e-7)
        q = net.res_gen.q_mvar.to_numpy()
        at_limit = np.isclose(np.abs(q), qlim, atol=1e-6)
        bound += int(at_limit.any())
        held = ~at_limit                                          # generators inside their limits hold their set-point
        np.testing.assert_allclose(net.res_gen.vm_pu.to_numpy()[held], net.gen.vm_pu.to_numpy()[held], atol=1e-9)
    assert (bound > 0) == expect_binding


@pytest.mark.parametrize("qlim,binding", [(1e4, False), (6.0, True)])
def test_per_environment_set_points_hostsim(qlim, binding):
    _check(qlim, binding, engine_cls=TorchHostSimEngine)


@pytest.mark.gpu
@pytest.mark.parametrize("qlim,binding", [(1e4, False), (6.0, True)])
def test_per_environment_set_points_cuda(cuda_lib, qlim, binding):
    _check(qlim, binding, sync=torch.cuda.synchronize)
