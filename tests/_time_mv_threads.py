"""Developer tool: LoadSheddingReconfiguration (111-bus MV stand-in with ties in the pattern -> k_pf_multi),
32 768 envs: step time by threads per environment."""
import sys; sys.path.insert(0, '.')
import torch
from opfgym_b200 import envs
B = 32768
for T in (int(x) for x in (sys.argv[1:] or ["32", "64"])):
    env = envs.LoadSheddingReconfiguration(num_envs=B, train_data="full_uniform", test_data="full_uniform", n_profile_steps=672,
                                           seed=1, copy_outputs=False, engine_kwargs=dict(threads_per_env=T))
    env.reset(seed=1)
    a = torch.rand(B, env.single_action_space.shape[0], dtype=torch.float64, device="cuda")
    for _ in range(4): env.step(a)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    K = 20
    for _ in range(K): env.step(a)
    e1.record(); torch.cuda.synchronize()
    i = env.engine.info
    print(f"T={T}: {e0.elapsed_time(e1)/K:.3f} ms/step  levels {i['n_levels']} blocks {i['n_blocks']} fill {i['n_fill_blocks']} smem/env {i['smem_bytes_pf']}", flush=True)
    env.close()
