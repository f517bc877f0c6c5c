"""Multi-stage episodes (reference opfgym/multi_stage.py:26-59)."""
import numpy as np
import pytest
import torch

from opfgym_b200 import envs
from opfgym_b200.multi_stage import MultiStageBatchedOpfEnv
from tests.hostsim.harness import TorchHostSimEngine


class MultiStageVC(MultiStageBatchedOpfEnv, envs.VoltageControl):
    pass


def make(n=8, k=3):
    return MultiStageVC(num_envs=n, steps_per_episode=k, engine_cls=TorchHostSimEngine,
                        n_profile_steps=4 * 672, obs_dtype="float64", train_data="simbench",
                        test_data="simbench", seed=1)


def test_episode_walks_consecutive_time_steps_and_resets():
    env = make()
    obs, _ = env.reset(seed=3)
    prof = torch.as_tensor(env.profiles[("load", "p_mw")].to_numpy().copy())
    start = env.current_simbench_step.clone()
    assert torch.allclose(env.col("load", "p_mw"), prof[start])
    act = torch.rand(8, 14, dtype=torch.float64)
    for k in (1, 2):
        q_before = None
        obs, reward, term, trunc, info = env.step(act)
        assert not term.any() and not trunc.any()
        assert torch.equal(env.current_simbench_step, start + k)
        assert torch.allclose(env.col("load", "p_mw"), prof[start + k])
        assert (env.step_in_episode == k).all()
    # like the reference, advancing runs `_sampling` only -- whose VoltageControl hook zeroes Q
    # (envs/voltage_control.py:133) -- and no centre action
    ctrl = env.positions("sgen", env.act_keys[0][2])
    assert (env.col("sgen", "q_mvar")[:, ctrl] == 0).all()
    obs, reward, term, trunc, info = env.step(act)
    assert term.all() and (env.step_in_episode == 0).all()                 # 3 steps -> terminated
    assert set(env.current_simbench_step.tolist()) <= set(env.train_steps.tolist())
    assert (env.col("sgen", "q_mvar")[:, ctrl] == 0).all()                 # fresh episode: centre action


def test_truncation_at_split_boundary():
    env = make(n=4, k=50)
    env.reset(seed=0, options={"step": 3 * 672 + 670})     # two steps before the end of the profiles
    act = torch.rand(4, 14, dtype=torch.float64)
    _, _, term, trunc, _ = env.step(act)
    assert not trunc.any() and not term.any()
    _, _, term, trunc, _ = env.step(act)
    assert trunc.all() and not term.any()                   # next step would leave the data
    assert (env.step_in_episode == 0).all()


def test_single_step_env_rejects_multi_step():
    with pytest.raises(NotImplementedError):
        envs.VoltageControl(num_envs=2, steps_per_episode=3, engine_cls=TorchHostSimEngine,
                            n_profile_steps=672, train_data="full_uniform", test_data="full_uniform")
