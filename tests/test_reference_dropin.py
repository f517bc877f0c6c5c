"""The drop-in boundary exercised by the reference's OWN code: the unmodified `opfgym` envs take this repo's
power flow through their constructor argument `power_flow_solver=` (opfgym/opf_env.py:53,70,657) and must behave
exactly as with their default power flow.  Needs /root/reference (build container only) -- the GPU box skips."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
needs_reference = pytest.mark.skipif(not os.path.isdir("/root/reference/opfgym"),
                                     reason="/root/reference is only present in the build container")


def _run(engine):
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "_ref_dropin_run.py"), engine],
                         capture_output=True, text=True, cwd=ROOT, timeout=900)
    assert res.returncode == 0, res.stderr[-2000:]
    return json.loads(res.stdout.strip().splitlines()[-1])


def _check(out):
    for name, worst in out.items():
        assert worst["steps"] == 3
        assert worst["obs"] < 1e-6, (name, worst)            # float32 observations
        assert worst["reward"] < 1e-8, (name, worst)
        assert worst["diverged_flag"] is False, name         # LoadflowNotConverged -> run_power_flow() == False


@needs_reference
def test_reference_env_with_plugged_power_flow_hostsim():
    _check(_run("hostsim"))


@needs_reference
@pytest.mark.gpu
def test_reference_env_with_plugged_power_flow_cuda(cuda_lib):
    _check(_run("cuda"))
