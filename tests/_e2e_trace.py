"""Developer tool (gpurun): CUPTI timeline of two step_host calls."""
import sys; sys.path.insert(0, '.')
import torch
from torch.profiler import profile, ProfilerActivity
from opfgym_b200 import envs

B = 32768
env = envs.VoltageControl(num_envs=B, train_data="full_uniform", test_data="full_uniform",
                          n_profile_steps=672, seed=1234, copy_outputs=False)
env.reset(seed=1)
h_act = torch.rand(B, 14, dtype=torch.float64).pin_memory()
mode = sys.argv[1] if len(sys.argv) > 1 else "host"
a_dev = h_act.cuda()
def one():
    if mode == "host":
        env.step_host(h_act)
    else:
        env.step(a_dev); torch.cuda.synchronize()
for _ in range(5):
    one()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        one()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t0 = evs[0].time_range.start
for e in evs:
    print(f"{(e.time_range.start - t0)/1e3:9.3f} ms  +{e.time_range.elapsed_us()/1e3:7.3f}  {e.name[:70]}")
