"""Callable constraints (reference constraints.py:27-64, examples/custom_constraint.py: an apparent-power limit of the
controllable generators on top of the default constraints) as BATCHED callables on the env: evaluated behind kernel
5, merged into validity, penalty, reward, cost and the info columns; checked per environment against the oracle's
restatement of the reference's reward arithmetic."""
import numpy as np
import pytest
import torch

from opfgym_b200 import constraints as C, grids
from opfgym_b200.opf_env import BatchedOpfEnv
from oracle import scoring
from tests.hostsim.harness import TorchHostSimEngine


def _env(n, limit, reward_function="summation", **kw):
    net, profiles = grids.build_simbench_net("1-MV-rural--0-sw", n_profile_steps=96)
    net.sgen["controllable"] = True
    net.sgen["min_q_mvar"], net.sgen["max_q_mvar"] = -0.3, 0.3
    obs_keys = [("load", "p_mw", net.load.index), ("load", "q_mvar", net.load.index), ("sgen", "p_mw", net.sgen.index)]
    act_keys = [("sgen", "q_mvar", net.sgen.index)]
    cons = C.create_default_constraints(net, {})
    n_default = len(cons)
    s_max = torch.as_tensor(net.sgen.max_max_p_mw.to_numpy(float) * limit)

    def s_mva(env):                                   # examples/custom_constraint.py:11-13, batched
        return (env.col("sgen", "p_mw") ** 2 + env.col("sgen", "q_mvar") ** 2) ** 0.5

    cons.append(C.Constraint("sgen", "s_mva", get_values=s_mva,
                             get_boundaries=lambda env: {"max": s_max.to(env.device)}, penalty_factor=2.0))
    env = BatchedOpfEnv(net, act_keys, obs_keys, profiles=profiles, num_envs=n, train_data="full_uniform",
                        test_data="full_uniform", seed=2, obs_dtype="float64", custom_constraints=cons,
                        reward_function=reward_function, **kw)
    return env, n_default, s_max.numpy()


def _check(reward_function, sync=lambda: None, **kw):
    n = 12
    env, n_default, s_max = _env(n, 0.45, reward_function, **kw)
    assert len(env.constraints) == n_default and len(env.batched_constraints) == 1
    env.reset(seed=3)
    act = torch.rand(n, env.single_action_space.shape[0], dtype=torch.float64, generator=torch.Generator().manual_seed(1))
    act[0] = 0.5                                      # q = 0: the least apparent power
    e = env.engine
    e.actions.copy_(act.to(env.device)); e.assemble(); sync()   # the cells this step will see
    p = env.col("sgen", "p_mw").cpu().numpy().copy()
    obs, reward, term, trunc, info = env.step(act.to(env.device))
    sync()
    assert info["valids"].shape == (n, n_default + 1) and info["violations"].shape == (n, n_default + 1)
    q_act = -0.3 + act.numpy() * 0.6
    s = np.sqrt(p ** 2 + q_act ** 2)
    excess = np.where(s > s_max, s - s_max, 0.0)
    viol = excess.sum(axis=1)
    np.testing.assert_allclose(info["violations"][:, -1].cpu().numpy(), viol, rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(info["unscaled_penalties"][:, -1].cpu().numpy(), -2.0 * viol, rtol=1e-12, atol=1e-15)
    assert (info["valids"][:, -1].cpu().numpy() == (excess.sum(axis=1) == 0)).all()
    assert 0 < int((viol > 0).sum()) <= n             # the limit binds somewhere
    # reward / cost: the reference's arithmetic over ALL constraints (oracle restatement of reward.py:61-98)
    objective = e.objective.cpu().numpy()
    pen = info["unscaled_penalties"].cpu().numpy().sum(axis=1)
    valid = info["valids"].cpu().numpy().all(axis=1)
    for b in range(n):
        want = scoring.reward(env.reward_function, float(objective[b]), float(pen[b]), bool(valid[b]))
        assert float(reward[b]) == pytest.approx(want, rel=1e-12, abs=1e-12)
        assert float(info["cost"][b]) == pytest.approx(scoring.cost(env.reward_function, float(pen[b]), bool(valid[b])),
                                                       rel=1e-12, abs=1e-12)
    # without the extra constraint the same step scores higher (or equal) wherever it is violated
    plain, _, _ = _env(n, 1e9, reward_function, **kw)
    plain.reset(seed=3)
    r0 = plain.step(act.to(plain.device))[1]
    sync()
    worse = (viol > 0)
    assert (reward.cpu().numpy()[worse] < r0.cpu().numpy()[worse]).all()
    np.testing.assert_allclose(reward.cpu().numpy()[~worse], r0.cpu().numpy()[~worse], rtol=1e-12)
    with pytest.raises(NotImplementedError):
        env.step_host(act.numpy())


@pytest.mark.parametrize("reward_function", ["summation", "replacement"])
def test_callable_constraint_hostsim(reward_function):
    _check(reward_function, engine_cls=TorchHostSimEngine)


@pytest.mark.gpu
def test_callable_constraint_cuda(cuda_lib):
    _check("summation", sync=torch.cuda.synchronize)


def test_getters_include_the_callable_constraint():
    env, n_default, s_max = _env(6, 0.45, engine_cls=TorchHostSimEngine)
    env.reset(seed=3)
    env._apply_actions(torch.rand(6, env.single_action_space.shape[0], dtype=torch.float64,
                                  generator=torch.Generator().manual_seed(1)))
    assert env.run_power_flow().all()
    valids, violations, penalties = env.calculate_violations()
    assert valids.shape == (6, n_default + 1) and violations.shape == (6, n_default + 1)
    s = (env.col("sgen", "p_mw") ** 2 + env.col("sgen", "q_mvar") ** 2) ** 0.5
    want = torch.clamp(s - torch.as_tensor(s_max), min=0.0).sum(dim=1)
    assert torch.allclose(violations[:, -1], want, rtol=1e-12, atol=1e-15)
    assert torch.equal(env.is_state_valid(), valids.all(dim=1))
