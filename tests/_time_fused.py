"""Developer tool (run under gpurun): time Engine.step() for the OPFG_FUSED_STEP setting of this process."""
import os, sys; sys.path.insert(0, '.')
import torch
from tests import common
from tests._time_quick import fill, timeit
from opfgym_b200.engine import Engine

name = sys.argv[1] if len(sys.argv) > 1 else "1-MV-semiurb--1-sw"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
case = common.make_case(name)
torch.manual_seed(0)
eng = Engine(case.program, B)
fill(case, eng, B)
ms = timeit(eng.step, n=30)
torch.cuda.synchronize()
r = eng.reward
print(f"fused={os.environ.get('OPFG_FUSED_STEP', '0')} {name} step={ms:.3f} ms  reward_sum={torch.nansum(r).item():.12e} "
      f"obs_sum={torch.nansum(eng.obs.double()).item():.12e} conv={eng.converged.sum().item()} launches/step={eng.launch_count()/35:.1f}")
