"""GPU tier for the "next"-list features that the CPU tier exercises on the host
build: multi-stage episodes, the mixed-grid batch, the StochasticObservation wrapper, samplers."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_multi_stage_cuda_equals_host_build(cuda_lib):
    import torch
    from opfgym_b200 import envs
    from opfgym_b200.multi_stage import MultiStageBatchedOpfEnv
    from tests.hostsim.harness import TorchHostSimEngine

    class MultiStageVC(MultiStageBatchedOpfEnv, envs.VoltageControl):
        pass

    kw = dict(num_envs=48, steps_per_episode=3, n_profile_steps=4 * 672, obs_dtype="float64",
              train_data="simbench", test_data="simbench", seed=1)
    gpu, cpu = MultiStageVC(**kw), MultiStageVC(engine_cls=TorchHostSimEngine, **kw)
    og, _ = gpu.reset(seed=3)
    oc, _ = cpu.reset(seed=3)
    np.testing.assert_allclose(og.cpu().numpy(), oc.numpy(), rtol=1e-13, atol=1e-15)
    for k in range(7):                         # two full episodes and the start of a third
        act = torch.rand(48, 14, dtype=torch.float64, generator=torch.Generator().manual_seed(k))
        rg, rc = gpu.step(act.cuda()), cpu.step(act)
        for i in (2, 3):                       # terminated, truncated
            assert torch.equal(rg[i].cpu(), rc[i])
        assert torch.equal(gpu.current_simbench_step.cpu(), cpu.current_simbench_step)
        np.testing.assert_allclose(rg[1].cpu().numpy(), rc[1].numpy(), rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(rg[0].cpu().numpy(), rc[0].numpy(), rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("fused", [True, False])
def test_mixed_batch_on_cuda_equals_members_alone(cuda_lib, fused):
    """BASELINE config 4: kernel 1 and kernel 5 of BOTH grids as one launch each (fused=True) -- every
    output bit equals the members stepped alone, and the launch counter shows the shared launches."""
    import torch
    from opfgym_b200 import envs
    from opfgym_b200.mixed import MixedBatchEnv
    kw = dict(n_profile_steps=672, obs_dtype="float64", train_data="full_uniform", test_data="full_uniform", seed=3)

    def members():
        return [envs.MaxRenewable(num_envs=27, **kw), envs.QMarket(num_envs=41, **kw)]
    mixed, alone = MixedBatchEnv(members(), fused_launches=fused), members()
    assert mixed.fused_launches == fused
    obs, _ = mixed.reset(seed=9)
    ref = [e.reset(seed=9)[0] for e in alone]
    assert torch.equal(obs[:27, :172], ref[0]) and torch.equal(obs[27:], ref[1])
    act = torch.rand(68, 18, dtype=torch.float64, device="cuda")
    eng = mixed.envs[0].engine
    for k in range(3):
        torch.cuda.synchronize()
        before = eng.launch_count()
        obs, reward, term, trunc, info = mixed.step(act)
        torch.cuda.synchronize()
        launches = eng.launch_count() - before
        r0, r1 = alone[0].step(act[:27, :18]), alone[1].step(act[27:, :10])
        assert torch.equal(reward, torch.cat([r0[1], r1[1]])) and term.all() and info["converged"].all()
        assert torch.equal(obs[:27, :172], r0[0]) and torch.equal(obs[27:], r1[0])
        assert torch.equal(info["cost"], torch.cat([r0[4]["cost"], r1[4]["cost"]]))
    assert mixed.episode_statistics()["steps"] == 68 * 3
    if fused:      # 1 assemble + 1 score for both grids (instead of 2 + 2); power flows and resets stay per member
        torch.cuda.synchronize()
        before = eng.launch_count()
        alone[0].step(act[:27, :18]); alone[1].step(act[27:, :10])
        torch.cuda.synchronize()
        assert launches == (eng.launch_count() - before) - 2


def test_wrapper_and_samplers_on_cuda(cuda_lib):
    import torch
    from opfgym_b200 import envs
    from opfgym_b200.wrappers import StochasticObservation
    kw = dict(n_profile_steps=4 * 672, obs_dtype="float64", seed=3)
    env = StochasticObservation(envs.QMarket(num_envs=64, train_data="noisy_simbench", test_data="simbench", **kw),
                                noise_relative_range=0.1)
    obs, _ = env.reset(seed=4)
    lo = torch.as_tensor(np.asarray(env.single_observation_space.low, float), device="cuda")
    hi = torch.as_tensor(np.asarray(env.single_observation_space.high, float), device="cuda")
    assert (obs >= lo - 1e-12).all() and (obs <= hi + 1e-12).all()
    obs, reward, term, trunc, info = env.step(torch.rand(64, 10, dtype=torch.float64, device="cuda"))
    assert term.all() and info["converged"].all() and not torch.isnan(reward).any()
    for data, params in (("normal_around_mean", {"relative_std": 0.2, "truncated": True}), ("mixed", {}),
                         ("noisy_simbench", {"noise_distribution": "normal", "interpolate_steps": True})):
        e = envs.VoltageControl(num_envs=256, train_data=data, test_data="simbench", sampling_params=params, **kw)
        e.reset(seed=1)
        _, reward, term, _, info = e.step(torch.rand(256, 14, dtype=torch.float64, device="cuda"))
        assert term.all() and info["converged"].float().mean() > 0.7, data
