"""Multi-stage episodes on the batched engine.

Reference ``opfgym/multi_stage.py:26-59``: an episode walks ``steps_per_episode``
consecutive SimBench time steps; after each ``step`` the next time step's state is
loaded (keeping the agent's set-points), the episode is *truncated* when the next
time step would leave the current data split and *terminated* after
``steps_per_episode`` steps.  Batched form: every environment carries its own
time-step index and step counter; finished environments are re-sampled while the
others advance, in the same launches (one profile-row gather for the whole batch).
"""
from __future__ import annotations

from .opf_env import BatchedOpfEnv


class MultiStageBatchedOpfEnv(BatchedOpfEnv):
    _multi_step_ok = True

    def __init__(self, *args, steps_per_episode: int = 4, **kwargs):
        assert steps_per_episode > 1, "At least two steps required for a multi-stage OPF."
        train_data = kwargs.get("train_data", "simbench")
        assert "simbench" in train_data, "multi-stage episodes need time-series (simbench) sampling"
        kwargs["prefetch_reset"] = False          # the next state depends on which envs finished
        super().__init__(*args, steps_per_episode=steps_per_episode, **kwargs)
        xp = self.xp
        self.step_in_episode = xp.zeros(self.num_envs, dtype=xp.int64, device=self.device)
        self._split_id = None

    def _split_of(self, steps):
        """0 = train, 1 = validation, 2 = test for every time-step index."""
        if self._split_id is None:
            import numpy as np
            from .data_split import N_STEPS
            ids = np.zeros(N_STEPS + 1, dtype=np.int64)
            ids[self.validation_steps.astype(np.int64)] = 1
            ids[self.test_steps.astype(np.int64)] = 2
            ids[N_STEPS] = -1
            self._split_id = self.engine._from_numpy(ids)
        return self._split_id[steps.clamp(max=self._split_id.shape[0] - 1)]

    def reset(self, seed=None, options=None):
        out = super().reset(seed=seed, options=options)
        self.step_in_episode.zero_()
        return out

    def step_host(self, actions=None):
        raise NotImplementedError("step_host pipelines single-step episodes; multi-stage episodes advance "
                                  "per environment -- use step() and copy what the host needs")

    def step(self, actions):
        xp, e = self.xp, self.engine
        act = xp.as_tensor(actions, device=self.device)
        e.actions.copy_(act.reshape(e.actions.shape))
        e.step(final_obs=True)
        self.power_flow_available = True
        self._results = e
        self.step_in_episode += 1
        nc = max(len(self.constraints), 1)
        info = {"valids": e.valids[:, :nc].bool(), "violations": e.violations[:, :nc].clone(),
                "unscaled_penalties": e.penalties[:, :nc].clone(), "cost": e.cost.clone(),
                "converged": e.converged.bool(), "iterations": e.iterations.clone(),
                "final_obs": self._obs_out(final=True).clone()}
        reward = e.reward.clone()
        cur = self.current_simbench_step
        nxt = cur + 1
        n_prof = len(next(iter(self.profiles.values())))
        # multi_stage.py:33-41: training never steps onto validation / test data, testing never onto
        # training data (the rule looks at the NEXT step only -- an episode started inside the other
        # split is truncated at once, as in the reference)
        nxt_split = self._split_of(nxt)
        truncated = ((nxt_split == 0) if self.test else (nxt_split > 0)) | (nxt >= n_prof)
        # the base class already flags the last stage as truncated (opf_env.py:408-410) before
        # multi_stage.py:43-45 adds terminated: both flags are set, as in the reference
        truncated = truncated | (self.step_in_episode >= self.steps_per_episode)
        terminated = (self.step_in_episode >= self.steps_per_episode) | ~info["converged"]   # :43-45, opf_env.py:399
        done = terminated | truncated
        # next state: finished envs draw a fresh time step (auto-reset), the others advance by one
        self._episode += 1
        self._stream_in_episode = 0
        pool = self._steps_dev[self.evaluate_on if self.test else "train"]
        u = xp.empty((self.num_envs, 1), dtype=xp.float64, device=self.device)
        e.philox_uniform(u, self.seed, self.first_env, self._next_stream())
        fresh = pool[(u[:, 0] * pool.shape[0]).long().clamp_(max=pool.shape[0] - 1)]
        slots = e.program.assembly["act_slot"]
        act_cols = xp.as_tensor(slots.astype("int64"), device=self.device)
        self._sampling(step=xp.where(done, fresh, nxt), test=self.test)
        kept = e.state[:, act_cols].clone()   # continuing envs: whatever _sampling (and its hook) left
        # reset applies the centre action to finished envs only (multi_stage.py:50 runs _sampling alone)
        if self.initial_action == "random":
            e.philox_uniform(e.actions_reset, self.seed, self.first_env, self._next_stream())
        else:
            e.actions_reset.fill_(0.5)
        e.assemble(scatter_sbus=False)
        e.state[:, act_cols] = xp.where(done[:, None], e.state[:, act_cols], kept)
        if self.pf_for_obs:
            e.assemble(apply_actions=False)
            self._reset_power_flow()
        else:
            e.observe()
        self.step_in_episode = xp.where(done, xp.zeros_like(self.step_in_episode), self.step_in_episode)
        return self._obs_out(), reward, terminated, truncated, info
