"""Developer tool: env.step with a host synchronisation after every step (an agent that looks at each result)."""
import sys, time; sys.path.insert(0, '.')
import torch
from opfgym_b200 import envs
B = 32768
env = envs.VoltageControl(num_envs=B, train_data="full_uniform", test_data="full_uniform", n_profile_steps=672, seed=1234,
                          copy_outputs=False)
env.reset(seed=1)
a = torch.rand(B, 14, dtype=torch.float64, device="cuda")
ts = []
for i in range(30):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    env.step(a); torch.cuda.synchronize()
    ts.append((time.perf_counter() - t0) * 1e3)
print("per-step wall with sync [ms]:", " ".join(f"{t:.2f}" for t in ts))
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(30): env.step(a)
torch.cuda.synchronize()
print(f"queued: {(time.perf_counter()-t0)/30*1e3:.3f} ms per step")
