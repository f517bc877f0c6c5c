"""Developer tool (gpurun): env.step throughput at small batch sizes, fused vs unfused reset."""
import sys; sys.path.insert(0, '.')
import torch
from opfgym_b200 import envs
from tests._time_quick import timeit
for B in (64, 512, 2048, 8192):
    for fused in (False, True):
        env = envs.VoltageControl(num_envs=B, train_data="full_uniform", test_data="full_uniform",
                                  n_profile_steps=672, seed=1, copy_outputs=False, fused_reset=fused)
        env.reset(seed=1)
        a = torch.rand(B, 14, dtype=torch.float64, device="cuda")
        ms = timeit(lambda: env.step(a), n=50, w=10)
        print(f"B={B:6d} fused={fused!s:5s} step={ms*1e3:8.1f} us  {B/ms*1e3:.3e} env-steps/s", flush=True)
        env.close(); del env
