"""Developer tool: time assemble and score kernels."""
import sys; sys.path.insert(0, '.')
import torch
from tests import common
from tests._time_quick import fill, timeit
from opfgym_b200.engine import Engine
case = common.make_case("1-MV-semiurb--1-sw")
B = 32768
eng = Engine(case.program, B)
fill(case, eng, B)
eng.step()
print(f"assemble {timeit(eng.assemble):.3f} ms  pf {timeit(eng.pf_solve):.3f} ms  score {timeit(eng.score):.3f} ms")
