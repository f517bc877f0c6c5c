"""Short run for ncu: a few full steps on the VoltageControl-size stand-in."""
import sys; sys.path.insert(0, '.')
import torch
from tests import common
from opfgym_b200.engine import Engine
name = sys.argv[1] if len(sys.argv) > 1 else "1-MV-semiurb--1-sw"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
n = int(sys.argv[3]) if len(sys.argv) > 3 else 3
case = common.make_case(name)
eng = Engine(case.program, B)
for t, c in common.SAMPLED:
    df = case.net[t]
    if len(df):
        lo = torch.tensor(df["min_min_" + c].to_numpy() / df.scaling.to_numpy(), device="cuda")
        hi = torch.tensor(df["max_max_" + c].to_numpy() / df.scaling.to_numpy(), device="cuda")
        eng.column(t, c).copy_(lo + (hi - lo) * torch.rand(B, len(df), device="cuda", dtype=torch.float64))
eng.actions.uniform_(0, 1)
for _ in range(n):
    eng.step()
torch.cuda.synchronize()
