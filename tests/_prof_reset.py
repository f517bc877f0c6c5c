"""Developer tool (gpurun, under ncu): a few fused episode resets at the bench size."""
import sys; sys.path.insert(0, '.')
import torch
from opfgym_b200 import envs
env = envs.VoltageControl(num_envs=32768, train_data="full_uniform", test_data="full_uniform",
                          n_profile_steps=672, seed=1234, copy_outputs=False, prefetch_reset=False)
env.reset(seed=1)
for _ in range(4):
    env._begin_episode()
torch.cuda.synchronize()
