"""Observation / action spaces.

Uses ``gymnasium.spaces.Box`` when gymnasium is importable; otherwise a minimal
duck-typed ``Box`` (gymnasium is absent from the build image).
``get_obs_and_state_space`` reproduces the bounds heuristics of reference
``opfgym/opf_env.py:720-803`` (row a10 of SURVEY.md §8): +-30 degree angles,
1.5x loading, +-0.75*band widening for voltages and ext-grid powers, bounds of
power columns divided by ``scaling``.
"""
from __future__ import annotations

import numpy as np

try:  # pragma: no cover - gymnasium is not installed in the build image
    from gymnasium.spaces import Box  # type: ignore
except Exception:  # noqa: BLE001
    class Box:  # minimal stand-in with gymnasium's float32 default
        def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
            if shape is None:
                shape = np.broadcast(np.asarray(low), np.asarray(high)).shape
            self.shape = tuple(shape)
            self.dtype = np.dtype(dtype)
            self.low = np.broadcast_to(np.asarray(low, dtype=float), self.shape).astype(self.dtype)
            self.high = np.broadcast_to(np.asarray(high, dtype=float), self.shape).astype(self.dtype)
            self._rng = np.random.default_rng(seed)

        def sample(self):
            return self._rng.uniform(self.low, self.high).astype(self.dtype)

        def seed(self, seed=None):
            self._rng = np.random.default_rng(seed)

        def contains(self, x):
            x = np.asarray(x)
            return bool(x.shape == self.shape and np.all(x >= self.low) and np.all(x <= self.high))

        def __repr__(self):
            return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"


def batch_space(space: Box, n: int) -> Box:
    return Box(np.tile(space.low, (n, 1)), np.tile(space.high, (n, 1)), dtype=space.dtype)


def get_obs_and_state_space(net, keys, add_time_obs=False, add_mean_obs=False, seed=None,
                            bus_wise_obs=False) -> Box:
    lows, highs = [], []
    if add_time_obs:
        lows.append(-np.ones(6))
        highs.append(np.ones(6))
    for unit_type, column, idxs in keys:
        if unit_type.startswith("res_"):
            unit_type = unit_type[4:]
        elif "max_" in column or "min_" in column:
            column = column[4:]
        table = net[unit_type]
        n = len(idxs)
        if column == "va_degree":
            lo, hi = np.full(n, -30.0), np.full(n, 30.0)
        else:
            def pick(prefix_wide, prefix):
                name = f"{prefix_wide}{column}" if f"{prefix_wide}{column}" in table.columns \
                    else f"{prefix}{column}"
                return table[name].loc[idxs].to_numpy(float)
            try:
                lo = pick("min_min_", "min_")
                hi = pick("max_max_", "max_")
            except KeyError:
                lo = np.zeros(n)
                hi = table[f"max_{column}"].loc[idxs].to_numpy(float) * 1.5
            if column == "vm_pu" or unit_type == "ext_grid":
                band = hi - lo
                lo, hi = lo - 0.75 * band, hi + 0.75 * band
        if "min" not in column and "max" not in column and "scaling" in table.columns:
            scal = table.scaling.loc[idxs].to_numpy(float)
            lo, hi = lo / scal, hi / scal
        if bus_wise_obs and unit_type == "load":       # opf_env.py:780-784
            at = table.bus.loc[idxs].to_numpy()
            buses = sorted(set(at.tolist()))
            lo = np.array([lo[at == b].sum() for b in buses])
            hi = np.array([hi[at == b].sum() for b in buses])
        if n > 0:
            lows.append(lo)
            highs.append(hi)
    if add_mean_obs:
        first = 1 if add_time_obs else 0
        lows.append(np.array([np.mean(l) for l in lows[first:] if len(l) > 1]))
        highs.append(np.array([np.mean(h) for h in highs[first:] if len(h) > 1]))
    low = np.concatenate(lows) if lows else np.zeros(0)
    high = np.concatenate(highs) if highs else np.zeros(0)
    assert not np.isnan(low).any() and not np.isnan(high).any(), "NaN in space bounds"
    return Box(low, high, seed=seed)
