"""Developer tool (gpurun): power-flow solve with and without the dense DC pre-pass (env OPFG_DC_PREPASS=0/1)."""
import sys; sys.path.insert(0, '.')
import torch
from tests import common
from tests._time_quick import fill, timeit
from opfgym_b200.engine import Engine
for name, B in (("1-MV-semiurb--1-sw", 32768), ("1-HV-urban--0-sw", 8192)):
    case = common.make_case(name)
    eng = Engine(case.program, B)
    fill(case, eng, B)
    eng.assemble()
    ms = timeit(eng.pf_solve)
    print(f"{name} pf_solve {ms:.3f} ms iters={eng.iterations.float().mean().item():.3f} conv={eng.converged.float().mean().item():.4f} "
          f"vm_sum={eng.vm.sum().item():.10f} launches/solve={eng.launch_count()/25:.1f}")
    eng.close()
