"""Small run of every kernel for compute-sanitizer (memcheck / racecheck)."""
import sys; sys.path.insert(0, ".")
import torch
from opfgym_b200 import envs
kw = dict(train_data="full_uniform", test_data="full_uniform", n_profile_steps=672)
for cls, n in ((envs.VoltageControl, 40), (envs.EcoDispatch, 12), (envs.LoadShedding, 24)):
    env = cls(num_envs=n, **kw)
    env.reset(seed=1)
    for _ in range(2):
        out = env.step(torch.rand(n, env.single_action_space.shape[0], device="cuda", dtype=torch.float64))
    torch.cuda.synchronize()
    assert out[4]["converged"].all()
    env.close()
    # unfused reset sequence, host-buffer step (three state buffers, side-stream look-ahead)
    env = cls(num_envs=n, fused_reset=False, **kw)
    env.reset(seed=2)
    for _ in range(3):
        out = env.step_host(torch.rand(n, env.single_action_space.shape[0]).numpy())
    assert out[4]["converged"].all()
    env.close()
print("sanitizer run ok")
