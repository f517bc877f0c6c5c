"""Developer tool: N-1 step time, all (environment, contingency) pairs as one batch vs one pass per contingency."""
import sys; sys.path.insert(0, '.')
import numpy as np, torch
from tests import test_islands
from opfgym_b200 import grids, net as pn, reward as R
from opfgym_b200.security_constrained import SecurityConstrainedBatchedOpfEnv

def make(n, batched, n_out):
    net, profiles = grids.build_simbench_net("1-MV-semiurb--1-sw", n_profile_steps=96)
    net.line = net.line[net.line.in_service.to_numpy(bool)]
    net.sgen["controllable"] = net.sgen.max_max_p_mw > np.sort(net.sgen.max_max_p_mw.to_numpy())[-9]
    net.sgen["min_p_mw"] = 0.0
    net.sgen["max_p_mw"] = net.sgen.max_max_p_mw
    outages = np.asarray(net.line.index[:: max(1, len(net.line) // n_out)][:n_out])
    return SecurityConstrainedBatchedOpfEnv(
        net, [("sgen", "p_mw", net.sgen.index[net.sgen.controllable])],
        [("load", "p_mw", net.load.index), ("sgen", "p_mw", net.sgen.index)], profiles=profiles, num_envs=n,
        train_data="full_uniform", test_data="full_uniform", seed=3,
        n_minus_one_keys=[("line", "in_service", outages)], reward_function=R.Summation(), batch_contingencies=batched)

for n, n_out in ((64, 20), (1024, 20), (8192, 20)):
    for batched in (True, False):
        env = make(n, batched, n_out)
        env.reset(seed=1)
        a = torch.rand(n, env.single_action_space.shape[0], dtype=torch.float64, device="cuda")
        for _ in range(3): env.step(a)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        K = 10
        for _ in range(K): out = env.step(a)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        print(f"envs {n:5d} x {n_out} contingencies, {'one batch' if batched else 'loop     '}: {ms:8.3f} ms/step "
              f"-> {n * (1 + n_out) / ms * 1e3:.3e} power flows/s", flush=True)
        env.close()
