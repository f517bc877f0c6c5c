"""N>1 path on CPU (gloo, world_size 2): env-index sharding with no data-path
collective, and the single all-reduce of episode statistics."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from opfgym_b200 import envs
    from tests.hostsim.harness import TorchHostSimEngine
    env = envs.VoltageControl(num_envs=6, rank=rank, world_size=world, engine_cls=TorchHostSimEngine,
                              train_data="full_uniform", test_data="full_uniform", seed=3,
                              n_profile_steps=672, obs_dtype="float64")
    obs, _ = env.reset(seed=8)
    env.reset_statistics()
    act = torch.rand(12, 14, dtype=torch.float64, generator=torch.Generator().manual_seed(0))
    _, reward, _, _, _ = env.step(act[rank * 6:(rank + 1) * 6])
    stats = env.episode_statistics(reduce=True)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), obs=obs.numpy(), reward=reward.numpy(),
             steps=stats["steps"], mean_reward=stats["mean_reward"])
    dist.destroy_process_group()


def test_two_rank_sharding_and_stats_allreduce(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = (np.load(tmp_path / f"rank{r}.npz") for r in (0, 1))
    # one process owning all 12 envs gives the same per-env results
    from opfgym_b200 import envs
    from tests.hostsim.harness import TorchHostSimEngine
    env = envs.VoltageControl(num_envs=12, engine_cls=TorchHostSimEngine, train_data="full_uniform",
                              test_data="full_uniform", seed=3, n_profile_steps=672, obs_dtype="float64")
    obs, _ = env.reset(seed=8)
    act = torch.rand(12, 14, dtype=torch.float64, generator=torch.Generator().manual_seed(0))
    _, reward, _, _, _ = env.step(act)
    np.testing.assert_array_equal(np.concatenate([r0["obs"], r1["obs"]]), obs.numpy())
    np.testing.assert_array_equal(np.concatenate([r0["reward"], r1["reward"]]), reward.numpy())
    # the all-reduced statistics are global and identical on both ranks
    assert r0["steps"] == r1["steps"] == 12
    assert np.isclose(r0["mean_reward"], reward.mean().item()) and r0["mean_reward"] == r1["mean_reward"]
