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
# the other power-flow kernels on the radial grid (auto = fused radial kernel above), per-environment
# Ybus values, per-environment voltage set-points, the profile sampler
from tests import common
from opfgym_b200.engine import Engine
case = common.make_case("1-MV-semiurb--1-sw")
for kernel in ("cta", "lanes", "radial"):
    eng = Engine(case.program, 70, obs_dtype="float64", pf_kernel=kernel, ordering=1)
    common.randomize(case, eng, seed=1)
    eng.step()
    torch.cuda.synchronize()
    assert eng.converged.all()
    eng.close()
env = envs.LoadSheddingReconfiguration(num_envs=24, **kw)
env.reset(seed=3)
env.step(torch.rand(24, env.single_action_space.shape[0], device="cuda", dtype=torch.float64))
env.close()
env = envs.VoltageControl(num_envs=33, train_data="noisy_simbench", test_data="simbench", n_profile_steps=4 * 672,
                          sampling_params=dict(noise_factor=0.1, interpolate_steps=True))
env.reset(seed=4)
env.step(torch.rand(33, 14, device="cuda", dtype=torch.float64))
torch.cuda.synchronize()
env.close()
# islands (kernel 1's connectivity walk, dropped buses in both power-flow kernels), switch cells with half-open
# lines and LV-side taps, N-1 with all contingencies as one batch
from tests import test_islands, test_switch_cells
for drop_ties in (True, False):
    env, _ = test_islands.make_n1_env(6, drop_ties=drop_ties)
    env.reset(seed=5)
    out = env.step(torch.rand(6, env.single_action_space.shape[0], device="cuda", dtype=torch.float64))
    assert out[4]["converged"].all()
    env.close()
env, sw = test_switch_cells.make_env(16, "lv")
env.reset(seed=6)
out = env.step(torch.rand(16, env.single_action_space.shape[0], device="cuda", dtype=torch.float64))
torch.cuda.synchronize()
assert torch.isnan(env.engine.vm).any() and out[4]["converged"].all()
env.close()
print("sanitizer run ok")
