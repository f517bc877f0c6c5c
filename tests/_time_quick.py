import sys, time; sys.path.insert(0,'.')
import numpy as np, torch
from tests import common
from opfgym_b200.engine import Engine
name = sys.argv[1] if len(sys.argv)>1 else "1-MV-semiurb--1-sw"
B = int(sys.argv[2]) if len(sys.argv)>2 else 32768
case = common.make_case(name)
eng = Engine(case.program, B)
print(eng.info)
for t, c in common.SAMPLED:
    df = case.net[t]
    if len(df):
        lo = torch.tensor(df["min_min_" + c].to_numpy() / df.scaling.to_numpy(), device="cuda")
        hi = torch.tensor(df["max_max_" + c].to_numpy() / df.scaling.to_numpy(), device="cuda")
        eng.column(t, c).copy_(lo + (hi - lo) * torch.rand(B, len(df), device="cuda", dtype=torch.float64))
eng.actions.uniform_(0, 1)
def timeit(fn, n=20, w=5):
    for _ in range(w): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n
for nm, fn in [('assemble', eng.assemble), ('pf', eng.pf_solve), ('score', eng.score), ('step', eng.step)]:
    ms = timeit(fn)
    print(f'{nm}: {ms:.3f} ms  -> {B/ms*1e3:.3e} env/s')
print('iters mean', eng.iterations.float().mean().item(), 'conv', eng.converged.float().mean().item())
