"""TEST INFRASTRUCTURE.  Host build of the engine sources for the GPU-less builder box.

``libopfg_hostsim.so`` is the same ``opfg_api.cu`` / ``opfg_core.h`` /
``symbolic.cpp`` compiled by g++ with ``-DOPFG_HOSTSIM``: "device memory" is
host memory and a "launch" is a loop over environments with one host thread per
environment (tid 0 of 1).  It lets the CPU test tier exercise the symbolic
schedule, the table compiler and the ctypes marshalling against the oracle.  It
cannot detect missing barriers or races -- that is what the ``-m gpu`` tier
(and compute-sanitizer) is for.  Nothing in ``opfgym_b200`` can load it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from opfgym_b200 import capi
from opfgym_b200.engine import Engine

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "..", "..", "opfgym_b200", "csrc")
LIB = os.path.join(HERE, "libopfg_hostsim.so")


def build(force=False):
    srcs = [os.path.join(CSRC, f) for f in ("opfg_api.cu", "symbolic.cpp")]
    deps = srcs + [os.path.join(CSRC, f) for f in ("opfg_core.h", "symbolic.hpp")] + \
        [os.path.join(HERE, "..", "..", "include", "opfg_b200.h")]
    if not force and os.path.exists(LIB) and all(
            os.path.getmtime(LIB) >= os.path.getmtime(d) for d in deps):
        return LIB
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-DOPFG_HOSTSIM", "-x", "c++",
           *srcs, "-o", LIB]
    subprocess.check_call(cmd)
    return LIB


def load():
    return capi.declare(C.CDLL(build()))


class HostSimEngine(Engine):
    """Engine plumbing on numpy arrays + the host-sim library (tests only)."""
    OPFG_TEST_ENGINE = True      # accepted by the env layer's engine_cls= test seam

    def __init__(self, program, num_envs, **kw):
        super().__init__(program, num_envs, lib=load(), **kw)

    def _setup_device(self, device):
        self.device = "host"

    def _zeros(self, shape, dtype):
        return np.zeros(shape, dtype=dtype)

    def _from_numpy(self, a):
        return np.ascontiguousarray(a).copy()

    def _ptr(self, t):
        return C.c_void_p(t.ctypes.data)

    def _stream(self):
        return C.c_void_p(0)


class TorchHostSimEngine(Engine):
    OPFG_TEST_ENGINE = True
    """Engine plumbing on torch CPU tensors + the host-sim library, so that the
    env layer (``BatchedOpfEnv`` and its tensor-op hooks) can be unit-tested on
    the GPU-less builder box.  Tests only."""

    def __init__(self, program, num_envs, device=None, **kw):
        super().__init__(program, num_envs, lib=load(), **kw)

    def _setup_device(self, device):
        import torch
        self.torch = torch
        self.device = torch.device("cpu")

    def _stream(self):
        return C.c_void_p(0)
