"""Developer tool (gpurun): device timeline of step_host calls -- when do the side-stream kernels of the look-ahead
episode finish, when does its observation transfer run, when does the step's own work end?"""
import sys; sys.path.insert(0, '.')
import torch
from opfgym_b200 import envs

B = 32768
env = envs.VoltageControl(num_envs=B, train_data="full_uniform", test_data="full_uniform",
                          n_profile_steps=672, seed=1234, copy_outputs=False)
env.reset(seed=1)
h_act = torch.rand(B, 14, dtype=torch.float64).pin_memory()
for _ in range(6):
    env.step_host(h_act)
torch.cuda.synchronize()
E = lambda: torch.cuda.Event(enable_timing=True)
log = []
main = torch.cuda.current_stream()
orig_begin = env._begin_episode
orig_step = env.engine.assemble
orig_score = env.engine.score
cur = {}

def begin(*a, **k):
    s = torch.cuda.current_stream()
    e0 = E(); e0.record(s)
    r = orig_begin(*a, **k)
    e1 = E(); e1.record(s)
    cur["side"] = (e0, e1)
    return r

def step(*a, **k):                       # step_host issues assemble / pf_solve / score separately
    cur["m0"] = E(); cur["m0"].record(main)
    return orig_step(*a, **k)


def score(*a, **k):
    r = orig_score(*a, **k)
    e1 = E(); e1.record(main)
    cur["main"] = (cur["m0"], e1)
    return r

env._begin_episode = begin
env.engine.assemble = step
env.engine.score = score
orig_copy = torch.Tensor.copy_
base = E(); base.record(main)
N = 8
for i in range(N):
    cur = {}
    env.step_host(h_act)
    # the observation transfer issued in this call: bracket it on the copy stream after the fact is not possible;
    # its end is the `ready` event of the pipe
    cur["ready"] = env._pipe["ready"]
    log.append(cur)
torch.cuda.synchronize()
end = E(); end.record(main); torch.cuda.synchronize()
print(f"{N} calls in {base.elapsed_time(end):.3f} ms -> {base.elapsed_time(end)/N:.3f} ms per call")
for i, c in enumerate(log):
    m0, m1 = c["main"]; s0, s1 = c["side"]
    print(f"call {i}: main {base.elapsed_time(m0):8.3f} .. {base.elapsed_time(m1):8.3f}   side kernels {base.elapsed_time(s0):8.3f} .. "
          f"{base.elapsed_time(s1):8.3f}   obs transfer of this call's look-ahead done at {base.elapsed_time(c['ready']):8.3f}")
