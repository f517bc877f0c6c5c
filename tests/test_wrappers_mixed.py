"""StochasticObservation wrapper (reference wrappers/stochastic_obs.py) and the mixed
MaxRenewable + QMarket batch (BASELINE config 4) on the host-sim build."""
import numpy as np
import torch

from opfgym_b200 import envs
from opfgym_b200.mixed import MixedBatchEnv
from opfgym_b200.wrappers import StochasticObservation
from tests.hostsim.harness import TorchHostSimEngine

KW = dict(engine_cls=TorchHostSimEngine, n_profile_steps=672, obs_dtype="float64",
          train_data="full_uniform", test_data="full_uniform", seed=3)


def test_stochastic_observation_noise_and_clipping():
    base = envs.QMarket(num_envs=16, **KW)
    clean, _ = base.reset(seed=4)
    env = StochasticObservation(envs.QMarket(num_envs=16, **KW), noise_relative_range=0.1)
    noisy, _ = env.reset(seed=4)
    lo = torch.as_tensor(np.asarray(env.single_observation_space.low, float))
    hi = torch.as_tensor(np.asarray(env.single_observation_space.high, float))
    assert noisy.shape == clean.shape and not torch.equal(noisy, clean)
    assert (noisy >= lo - 1e-12).all() and (noisy <= hi + 1e-12).all()       # clipped to the space
    assert ((noisy - clean).abs() <= 0.1 * (hi - lo) + 1e-9).all()
    wide = StochasticObservation(envs.QMarket(num_envs=16, **KW), noise_relative_range=0.1,
                                 maintain_original_range=False)
    assert (np.asarray(wide.single_observation_space.high) >
            np.asarray(base.single_observation_space.high) - 1e-12).all()
    obs, reward, term, trunc, info = env.step(torch.rand(16, 10, dtype=torch.float64))
    assert obs.shape == (16, 305) and term.all()
    assert env.num_envs == 16 and env.single_action_space.shape == (10,)      # attribute pass-through


import pytest


@pytest.mark.parametrize("fused", [True, False])
def test_mixed_batch_equals_members_stepped_alone(fused):
    def members():
        return [envs.MaxRenewable(num_envs=6, **KW), envs.QMarket(num_envs=10, **KW)]
    mixed = MixedBatchEnv(members(), fused_launches=fused)
    assert mixed.fused_launches == fused
    alone = members()
    assert mixed.num_envs == 16 and mixed.n_obs == 305 and mixed.n_act == 18
    obs, _ = mixed.reset(seed=9)
    ref = [e.reset(seed=9)[0] for e in alone]
    assert torch.equal(obs[:6, :172], ref[0]) and torch.isnan(obs[:6, 172:]).all()
    assert torch.equal(obs[6:], ref[1])
    act = torch.rand(16, 18, dtype=torch.float64, generator=torch.Generator().manual_seed(0))
    obs, reward, term, trunc, info = mixed.step(act)
    r0 = alone[0].step(act[:6, :18])
    r1 = alone[1].step(act[6:, :10])
    assert torch.equal(reward, torch.cat([r0[1], r1[1]])) and term.all() and info["converged"].all()
    assert torch.equal(obs[:6, :172], r0[0]) and torch.equal(obs[6:], r1[0])
    assert torch.equal(info["cost"], torch.cat([r0[4]["cost"], r1[4]["cost"]]))
    assert mixed.episode_statistics()["steps"] == 16
    obs2, reward2 = mixed.step(act)[:2]                 # second step: buffers were switched
    assert torch.equal(reward2, torch.cat([alone[0].step(act[:6, :18])[1], alone[1].step(act[6:, :10])[1]]))


def _host_step_equals_device_step(device_kw, sync=lambda: None):
    def members():
        return [envs.MaxRenewable(num_envs=6, **device_kw), envs.QMarket(num_envs=10, **device_kw)]
    a, b = MixedBatchEnv(members()), MixedBatchEnv(members())
    a.reset(seed=9); b.reset(seed=9)
    for k in range(3):                                  # the pinned result buffers are reused from the second call on
        act = torch.rand(16, 18, dtype=torch.float64, generator=torch.Generator().manual_seed(k))
        obs, reward, term, trunc, info = a.step(act.to(a.device))
        sync()
        h_obs, h_reward, h_term, h_trunc, h_info = b.step_host(act.numpy())
        assert isinstance(h_obs, np.ndarray) and h_obs.shape == (16, 305)
        np.testing.assert_array_equal(h_obs, obs.cpu().numpy())           # NaN padding compares equal here
        np.testing.assert_array_equal(h_reward, reward.cpu().numpy())
        np.testing.assert_array_equal(h_info["cost"], info["cost"].cpu().numpy())
        assert h_term.all() and h_info["converged"].all() and not h_trunc.any()


def test_mixed_step_host_equals_step():
    _host_step_equals_device_step(KW)


@pytest.mark.gpu
def test_mixed_step_host_equals_step_cuda(cuda_lib):
    kw = {k: v for k, v in KW.items() if k != "engine_cls"}
    _host_step_equals_device_step(kw, sync=torch.cuda.synchronize)
