"""Developer tool: the meshed-grid power-flow kernel (shared block slots, gathers beside the diagonal phase) for
compute-sanitizer racecheck / memcheck -- a short run of the three grids that take it."""
import sys; sys.path.insert(0, ".")
import torch
from opfgym_b200 import envs
kw = dict(train_data="full_uniform", test_data="full_uniform", n_profile_steps=96)
for cls, n in ((envs.EcoDispatch, 7), (envs.MaxRenewable, 5), (envs.LoadSheddingReconfiguration, 26)):
    env = cls(num_envs=n, **kw)
    env.reset(seed=1)
    out = env.step(torch.rand(n, env.single_action_space.shape[0], device="cuda", dtype=torch.float64))
    torch.cuda.synchronize()
    assert out[4]["converged"].all()
    print(cls.__name__, env.engine.info["smem_bytes_pf"], "B per environment, blocks", env.engine.info["n_blocks"], flush=True)
    env.close()
