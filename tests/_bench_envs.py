"""Developer tool: device-resident throughput of the five benchmark envs and the mixed batch."""
import sys; sys.path.insert(0, ".")
import torch
from opfgym_b200 import envs
from opfgym_b200.mixed import MixedBatchEnv

kw = dict(train_data="full_uniform", test_data="full_uniform", n_profile_steps=672)


def bench(env, n_act, steps=20, warm=4, B=None):
    B = B or env.num_envs
    act = torch.rand(B, n_act, device="cuda", dtype=torch.float64)
    env.reset(seed=1)
    for _ in range(warm):
        env.step(act)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        env.step(act)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return ms, B / ms * 1e3


for name, cls, B in (("VoltageControl", envs.VoltageControl, 32768), ("QMarket", envs.QMarket, 32768),
                     ("LoadShedding", envs.LoadShedding, 32768), ("EcoDispatch", envs.EcoDispatch, 8192),
                     ("MaxRenewable", envs.MaxRenewable, 8192)):
    env = cls(num_envs=B, copy_outputs=False, **kw)
    env.reset_statistics()
    ms, rate = bench(env, env.single_action_space.shape[0])
    s = env.episode_statistics()
    i = env.engine.info
    print(f"{name:15s} nb={i['nb']:4d} B={B:6d} {ms:7.3f} ms/step {rate:.3e} env-steps/s  conv={s['converged_share']:.4f} "
          f"valid={s['valid_share']:.3f} iters={s['mean_iterations']:.2f} levels={i['n_levels']} T={i['threads_per_env']}", flush=True)
    env.close()
    del env
mix = MixedBatchEnv([envs.MaxRenewable(num_envs=8192, copy_outputs=False, **kw),
                     envs.QMarket(num_envs=24576, copy_outputs=False, **kw)])
ms, rate = bench(mix, mix.n_act, B=mix.num_envs)
print(f"Mixed MR+QM     B={mix.num_envs:6d} {ms:7.3f} ms/step {rate:.3e} env-steps/s")
