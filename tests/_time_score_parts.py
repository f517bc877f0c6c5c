"""Developer tool (gpurun): kernel 5 with parts switched off through NULL batch pointers."""
import sys; sys.path.insert(0, '.')
import ctypes as C
import torch
from tests import common
from tests._time_quick import fill, timeit
from opfgym_b200 import capi
from opfgym_b200.engine import Engine
case = common.make_case("1-MV-semiurb--1-sw")
B = 32768
eng = Engine(case.program, B)
fill(case, eng, B)
eng.assemble(); eng.pf_solve()
def run(batch):
    capi.check(eng.lib, eng.lib.opfg_score(eng.handle, C.byref(batch), eng._stream()))
full = capi.Batch.from_buffer_copy(eng.batch)
print(f"full               {timeit(lambda: run(full))*1e3:7.1f} us")
b = capi.Batch.from_buffer_copy(eng.batch); b.obs_f32 = None; b.obs_f64 = None
print(f"no observation     {timeit(lambda: run(b))*1e3:7.1f} us")
b2 = capi.Batch.from_buffer_copy(b); b2.stats = None
print(f"... no stats       {timeit(lambda: run(b2))*1e3:7.1f} us")
b3 = capi.Batch.from_buffer_copy(b2); b3.valids = None; b3.violations = None; b3.penalties = None
print(f"... no info arrays {timeit(lambda: run(b3))*1e3:7.1f} us")
