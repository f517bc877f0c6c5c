"""Stochastic observations for the batched env.

Reference ``opfgym/wrappers/stochastic_obs.py:10-52``: every observation gets
uniform noise of ``noise_relative_range`` x (space high - low) and is optionally
clipped back into the original space.  Here the noise comes from the engine's
counter-based generator (``opfg_philox_uniform``, keyed by the global env id, so
it does not depend on how envs are sharded over GPUs) and is applied to the whole
``[num_envs, n_obs]`` batch with two tensor ops.
"""
from __future__ import annotations

import numpy as np

from ..spaces import Box, batch_space


class StochasticObservation:
    def __init__(self, env, noise_relative_range: float = 0.1, maintain_original_range: bool = True,
                 seed: int = 0):
        self.env = env
        self.maintain_original_range = maintain_original_range
        space = env.single_observation_space
        low, high = np.asarray(space.low, float), np.asarray(space.high, float)
        self.abs_noise_range = noise_relative_range * (high - low)
        if not maintain_original_range:
            self.single_observation_space = Box(low - self.abs_noise_range, high + self.abs_noise_range)
        else:
            self.single_observation_space = space
        self.observation_space = batch_space(self.single_observation_space, env.num_envs)
        xp = env.xp
        self._range = xp.as_tensor(self.abs_noise_range, device=env.device)
        self._low = xp.as_tensor(low, device=env.device)
        self._high = xp.as_tensor(high, device=env.device)
        self._u = xp.empty((env.num_envs, len(low)), dtype=xp.float64, device=env.device)
        self._seed, self._calls = int(seed), 0

    def __getattr__(self, name):
        return getattr(self.env, name)

    def observation(self, obs):
        """stochastic_obs.py:40-52: obs + U(-r, r), clipped if ``maintain_original_range``."""
        self._calls += 1
        self.env.engine.philox_uniform(self._u, self._seed ^ 0x5EED, self.env.first_env,
                                       (1 << 40) + self._calls)
        noisy = obs.to(self._u.dtype) + (self._u * 2.0 - 1.0) * self._range
        if self.maintain_original_range:
            noisy = self.env.xp.minimum(self.env.xp.maximum(noisy, self._low), self._high)
        return noisy.to(obs.dtype)

    def reset(self, seed=None, options=None):
        if seed is not None:
            self._seed, self._calls = int(seed), 0
        obs, info = self.env.reset(seed=seed, options=options)
        return self.observation(obs), info

    def step(self, actions):
        obs, reward, terminated, truncated, info = self.env.step(actions)
        return self.observation(obs), reward, terminated, truncated, info
