"""Developer tool (gpurun): where does a step_host call spend its time?"""
import sys, time; sys.path.insert(0, '.')
import torch
from opfgym_b200 import envs

B = 32768
env = envs.VoltageControl(num_envs=B, train_data="full_uniform", test_data="full_uniform",
                          n_profile_steps=672, seed=1234, copy_outputs=False)
env.reset(seed=1)
h_act = torch.rand(B, 14, dtype=torch.float64).pin_memory()
for _ in range(5):
    env.step_host(h_act)
# host issue time vs total
orig_sync = torch.cuda.Stream.synchronize
marks = {}
def patched(self):
    marks['issue'] = time.perf_counter()
    return orig_sync(self)
torch.cuda.Stream.synchronize = patched
tot = iss = 0
N = 20
for _ in range(N):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    env.step_host(h_act)
    t1 = time.perf_counter()
    tot += t1 - t0; iss += marks['issue'] - t0
print(f"step_host wall {tot/N*1e3:.3f} ms, host issue part {iss/N*1e3:.3f} ms")
torch.cuda.Stream.synchronize = orig_sync
# raw copy speeds
obs = env.engine.obs
h = env._host["obs"]
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): h.copy_(obs, non_blocking=True)
e1.record(); torch.cuda.synchronize()
print(f"obs D2H {e0.elapsed_time(e1)/10:.3f} ms for {obs.numel()*obs.element_size()/1e6:.1f} MB")
# device-only step with sync each step (no host copies)
env.reset(seed=2)
a = h_act.cuda()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(N):
    env.step(a); torch.cuda.synchronize()
print(f"env.step + sync wall {(time.perf_counter()-t0)/N*1e3:.3f} ms")
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(N):
    env.step(a)
torch.cuda.synchronize()
print(f"env.step queued wall {(time.perf_counter()-t0)/N*1e3:.3f} ms")
