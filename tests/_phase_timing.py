"""Developer tool: per-phase cycle breakdown of k_pf (needs the -DOPFG_PHASE_TIMING build)."""
import sys; sys.path.insert(0, '.')
import ctypes as C
import numpy as np, torch
from tests import common
from opfgym_b200 import capi
from opfgym_b200.engine import Engine

name = sys.argv[1] if len(sys.argv) > 1 else "1-MV-semiurb--1-sw"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
T = int(sys.argv[3]) if len(sys.argv) > 3 else 64
lib = capi.declare(C.CDLL("opfgym_b200/lib/libopfg_b200_timing.so"))
lib.opfg_debug_phase_cycles.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
case = common.make_case(name)
eng = Engine(case.program, B, lib=lib, threads_per_env=T)
for t, c in common.SAMPLED:
    df = case.net[t]
    if len(df):
        lo = torch.tensor(df["min_min_" + c].to_numpy() / df.scaling.to_numpy(), device="cuda")
        hi = torch.tensor(df["max_max_" + c].to_numpy() / df.scaling.to_numpy(), device="cuda")
        eng.column(t, c).copy_(lo + (hi - lo) * torch.rand(B, len(df), device="cuda", dtype=torch.float64))
eng.actions.uniform_(0, 1)
eng.assemble()
lib.opfg_debug_phase_cycles(eng.handle, None, 1)
eng.pf_solve(); eng.pf_solve()
lib.opfg_debug_phase_cycles(eng.handle, None, 1)
n = 5
for _ in range(n): eng.pf_solve()
out = np.zeros(144, np.uint64)
lib.opfg_debug_phase_cycles(eng.handle, out.ctypes.data, 1)
names = ["dc+init", "rows+J", "rows only", "lu diag", "lu off", "bwd", "update", "output"]
per_env = out.astype(float) / (n * B)
print(f"{name} T={T} cycles per env (thread 0): total {per_env.sum():.0f}")
for k, v in zip(names, per_env[:8]):
    print(f"  {k:10s} {v:9.0f}  {100*v/per_env.sum():5.1f}%")
nl = eng.info["n_levels"]
it = float(eng.iterations.float().mean())
print("per level, cycles per NR iteration: lane 0 diagonal + rest of the phase (gathers, barrier) / scale / bwd")
for l in range(nl):
    print(f"  L{l:<2d} {per_env[112+l]/it:8.0f} + {per_env[16+l]/it:8.0f} {per_env[48+l]/it:8.0f} {per_env[80+l]/it:8.0f}")
