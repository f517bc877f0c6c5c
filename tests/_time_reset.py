"""Developer tool (gpurun): cost of the fused reset kernel by stage subset."""
import sys; sys.path.insert(0, '.')
import torch
from opfgym_b200 import envs
from tests._time_quick import timeit

B = 32768
env = envs.VoltageControl(num_envs=B, train_data="full_uniform", test_data="full_uniform",
                          n_profile_steps=672, seed=1234, copy_outputs=False, prefetch_reset=False)
env.reset(seed=1)
e = env.engine
# re-record the trace
env._reset_plans.clear()
e.trace = []
env._sampling(None, False, True)
trace, e.trace = e.trace, None
print([t[0] for t in trace], "n_inputs", env.program.layout.n_inputs, "n_state", env.program.layout.n)
for name, sub in (("none", []), ("sample", trace[:1]), ("sample+sgen", trace[:2]), ("all", trace)):
    plan = e.make_reset_plan(sub, 0)
    ms = timeit(lambda: e.reset_episode(plan, 1, 0, 64, False, 9), n=20)
    print(f"{name:12s} {ms*1e3:8.1f} us")
def seq():
    env.fused_reset = False
    env._begin_episode()
ms = timeit(seq, n=20)
print(f"unfused sequence {ms*1e3:8.1f} us")
