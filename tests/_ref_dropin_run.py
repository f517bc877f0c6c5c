"""Helper of tests/test_reference_dropin.py (run as a subprocess: it installs stub modules for pandapower /
simbench / gymnasium, which must not leak into the test session).

The UNMODIFIED reference env (/root/reference/opfgym, `VoltageControl` and `EcoDispatch`) is stepped twice with
the same seeds: once with its default power flow (under the stubs `pp.runpp` = the CPU oracle) and once with this
repo's plug-in handed to the reference's own constructor argument `power_flow_solver=`
(opfgym/opf_env.py:53,70,657).  Prints the largest differences as JSON."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import _ref_stubs                                                   # noqa: E402

_ref_stubs.install(profile_steps=96)
import opfgym.envs as ref_envs                                      # noqa: E402  (the reference, unmodified)

from opfgym_b200 import adapter                                     # noqa: E402

engine = sys.argv[1] if len(sys.argv) > 1 else "hostsim"
kw = {}
if engine == "hostsim":
    from tests.hostsim.harness import TorchHostSimEngine
    kw["engine_cls"] = TorchHostSimEngine

out = {}
for name in ("VoltageControl", "EcoDispatch"):
    cls = getattr(ref_envs, name)
    plain = cls(train_data="full_uniform", test_data="full_uniform", seed=7)
    plugged = cls(train_data="full_uniform", test_data="full_uniform", seed=7,
                  power_flow_solver=adapter.PowerFlowSolver(plain.net, **kw))
    worst = dict(obs=0.0, reward=0.0, steps=0, solver_calls=0)
    for episode in range(3):
        o0, _ = plain.reset(seed=100 + episode)
        o1, _ = plugged.reset(seed=100 + episode)
        worst["obs"] = max(worst["obs"], float(np.nanmax(np.abs(o0 - o1))))
        act = np.random.default_rng(episode).uniform(0, 1, plain.action_space.shape).astype(np.float32)
        r0 = plain.step(act)
        r1 = plugged.step(act)
        worst["obs"] = max(worst["obs"], float(np.nanmax(np.abs(r0[0] - r1[0]))))
        worst["reward"] = max(worst["reward"], abs(float(r0[1]) - float(r1[1])))
        assert r0[2] == r1[2] and r0[3] == r1[3]
        for key in ("valids", "violations", "unscaled_penalties"):
            np.testing.assert_allclose(np.asarray(r0[4][key], float), np.asarray(r1[4][key], float), atol=1e-7)
        worst["steps"] += 1
    # the reference's own flag mapping: a diverging state makes run_power_flow() return False (opf_env.py:656-662)
    plugged.net.load["p_mw"] *= 500.0
    worst["diverged_flag"] = bool(plugged.run_power_flow())
    out[name] = worst
print(json.dumps(out))
