"""Developer tool (gpurun): cost of the unfused next-episode kernels (sampler, hook programs, observe)."""
import sys; sys.path.insert(0, '.')
import torch
from opfgym_b200 import envs
from tests._time_quick import timeit
B = 32768
env = envs.VoltageControl(num_envs=B, train_data="full_uniform", test_data="full_uniform",
                          n_profile_steps=672, seed=1234, copy_outputs=False, prefetch_reset=False, fused_reset=False)
env.reset(seed=1)
e = env.engine
print(f"begin_episode {timeit(env._begin_episode, n=30)*1e3:7.1f} us", end="  ")
print(f"sampler {timeit(lambda: env._sample_uniform(), n=30)*1e3:7.1f} us", end="  ")
for key, prog in env._row_programs.items():
    if prog is not None:
        print(f"{key} {timeit(prog.run, n=30)*1e3:7.1f} us", end="  ")
print(f"observe {timeit(e.observe, n=30)*1e3:7.1f} us")
