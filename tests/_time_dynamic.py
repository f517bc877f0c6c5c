"""Developer tool (gpurun): BASELINE config 5 -- tap / switch actions (per-environment Ybus values)."""
import sys; sys.path.insert(0, '.')
import torch
from tests.test_dynamic_branches import make_env
from tests._time_quick import timeit
B = 32768
env, ties = make_env(B, copy_outputs=False)
env.reset(seed=1)
n_act = env.single_action_space.shape[0]
a = torch.rand(B, n_act, dtype=torch.float64, device="cuda")
ms = timeit(lambda: env.step(a), n=20)
e = env.engine
print(f"config 5: nb={e.info['nb']} levels={e.info['n_levels']} n_act={n_act} step={ms:.3f} ms -> {B/ms*1e3:.3e} env-steps/s "
      f"conv={e.converged.float().mean().item():.4f} iters={e.iterations.float().mean().item():.2f}")
print(f"  assemble {timeit(e.assemble)*1e3:.0f} us  pf {timeit(e.pf_solve)*1e3:.0f} us  score {timeit(e.score)*1e3:.0f} us")
