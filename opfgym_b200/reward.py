"""Reward-function *descriptions* for the batched engine.

Same class names, constructor arguments and scaling helpers as reference
``opfgym/reward.py`` (``RewardFunction`` :8-106, scalers :120-178,
``Summation`` :219, ``Replacement`` :230, ``Parameterized`` :254,
``OnlyObjective`` :308).  The arithmetic of ``__call__``/``calculate_cost``
 runs in the epilogue of the fused scoring kernel for all
environments at once; the methods below are the *scalar host mirror* of that
arithmetic (used for configuration, for `estimate`-style tooling and by the
unit tests that replay the reference's own known-answer cases).
"""
from __future__ import annotations

import math

import numpy as np

SUMMATION, REPLACEMENT, PARAMETERIZED, ONLY_OBJECTIVE = 0, 1, 2, 3


def calculate_normalization_params(std_objective, mean_objective, std_penalty,
                                   mean_penalty, **_):
    """(x - mean) / std  ->  factor, bias  (reference :120-137)."""
    return {"objective_factor": 1.0 / std_objective,
            "objective_bias": -mean_objective / std_objective,
            "penalty_factor": 1.0 / std_penalty,
            "penalty_bias": -mean_penalty / std_penalty}


def calculate_minmax01_params(min_objective, max_objective, min_penalty,
                              max_penalty, **_):
    """[min, max] -> [0, 1]  (reference :140-158)."""
    d_obj = max_objective - min_objective
    d_pen = max_penalty - min_penalty
    return {"objective_factor": 1.0 / d_obj, "objective_bias": -(min_objective / d_obj),
            "penalty_factor": 1.0 / d_pen, "penalty_bias": -(min_penalty / d_pen)}


def calculate_minmax11_params(min_objective, max_objective, min_penalty,
                              max_penalty, **_):
    """[min, max] -> [-1, 1]  (reference :161-178)."""
    h_obj = (max_objective - min_objective) / 2.0
    h_pen = (max_penalty - min_penalty) / 2.0
    return {"objective_factor": 1.0 / h_obj, "objective_bias": -(min_objective / h_obj + 1.0),
            "penalty_factor": 1.0 / h_pen, "penalty_bias": -(min_penalty / h_pen + 1.0)}


_SCALERS = {"minmax11": calculate_minmax11_params,
            "minmax01": calculate_minmax01_params,
            "normalization": calculate_normalization_params}


def select_reward_scaler(reward_scaling: str):
    try:
        return _SCALERS[reward_scaling]
    except KeyError:
        raise NotImplementedError("This reward scaling does not exist!") from None


def estimate_reward_distribution(env, num_samples: int = 3000) -> dict:
    """Reference :181-216 loops ``num_samples`` x (reset, random action, power
    flow) on one env.  Here the whole sample is ONE batched engine call."""
    objectives, penalties = env.sample_objective_penalty(num_samples)
    objectives = objectives[~np.isnan(objectives)]
    penalties = penalties[~np.isnan(penalties)]
    out = {}
    for name, arr in (("objective", objectives), ("penalty", penalties)):
        out["min_" + name] = arr.min()
        out["max_" + name] = arr.max()
        out["mean_" + name] = arr.mean()
        out["std_" + name] = np.std(arr)
        out["median_" + name] = np.median(arr)
        out["mean_abs_" + name] = np.abs(arr).mean()
    return out


class RewardFunction:
    kind = SUMMATION

    def __init__(self, penalty_weight: float | None = 0.5,
                 clip_range: tuple[float, float] | None = None,
                 reward_scaling: str | None = None,
                 scaling_params: dict | None = None,
                 env=None):
        self.penalty_weight = penalty_weight
        self.clip_range = clip_range
        self.scaling_params = self.prepare_reward_scaling(reward_scaling, scaling_params, env)

    def prepare_reward_scaling(self, reward_scaling, scaling_params, env) -> dict:
        if not isinstance(reward_scaling, str):
            return {"penalty_factor": 1, "penalty_bias": 0,
                    "objective_factor": 1, "objective_bias": 0}
        params = dict(scaling_params or {})
        user = dict(params)
        scaler = select_reward_scaler(reward_scaling)
        try:
            params.update(scaler(**params))
        except TypeError:
            params = estimate_reward_distribution(env, **params)
            params.update(scaler(**params))
        params.update(user)
        if np.isnan(params["penalty_bias"]):
            params["penalty_bias"] = 0
        if np.isinf(params["penalty_factor"]):
            params["penalty_factor"] = 1
        return params

    # ---- scalar host mirror of the kernel epilogue -------------------------
    def __call__(self, objective: float, penalty: float, valid: bool) -> float:
        objective = self.scale_objective(self.adjust_objective(objective, valid))
        penalty = self.scale_penalty(self.adjust_penalty(penalty, valid))
        reward = self.compute_total_reward(objective, penalty)
        return self.clip_reward(reward) if self.clip_range else reward

    def clip_reward(self, reward: float) -> float:
        lo, hi = self.clip_range
        return min(max(reward, lo), hi)

    def compute_total_reward(self, objective: float, penalty: float) -> float:
        w = self.penalty_weight
        if w is None:
            return objective + penalty
        return objective * (1 - w) + penalty * w

    def scale_objective(self, objective: float) -> float:
        return objective * self.scaling_params["objective_factor"] + self.scaling_params["objective_bias"]

    def scale_penalty(self, penalty: float) -> float:
        return penalty * self.scaling_params["penalty_factor"] + self.scaling_params["penalty_bias"]

    def calculate_cost(self, penalty: float, valid: bool) -> float:
        return 0.0 if valid else abs(penalty * self.scaling_params["penalty_factor"])

    def adjust_penalty(self, penalty: float, valid: bool) -> float:
        return penalty

    def adjust_objective(self, objective: float, valid: bool) -> float:
        return objective

    # ---- tensor form of the same arithmetic (N-1 recombination on the device) ----
    def batched(self, objective, penalty, valid):
        """``__call__`` on torch tensors [B] (objective, penalty float; valid bool)."""
        import torch
        p = self.device_params()
        obj, pen = objective, penalty
        if p["kind"] == REPLACEMENT:
            obj = torch.where(valid, obj + p["valid_reward"], torch.zeros_like(obj))
        elif p["kind"] == PARAMETERIZED:
            pen = torch.where(valid, pen + p["valid_reward"], pen - p["invalid_penalty"])
            obj = torch.where(valid, obj, obj * p["invalid_objective_share"])
        elif p["kind"] == ONLY_OBJECTIVE:
            pen = torch.zeros_like(pen)
        obj = obj * p["objective_factor"] + p["objective_bias"]
        pen = pen * p["penalty_factor"] + p["penalty_bias"]
        w = p["penalty_weight"]
        r = obj + pen if math.isnan(w) else obj * (1 - w) + pen * w
        if not math.isnan(p["clip_lo"]):
            r = r.clamp(p["clip_lo"], p["clip_hi"])
        return r

    def batched_cost(self, penalty, valid):
        import torch
        p = self.device_params()
        c = (penalty * p["penalty_factor"]).abs()
        if p["kind"] == PARAMETERIZED:
            c = c + p["invalid_penalty"]
        return torch.where(valid, torch.zeros_like(c), c)

    # ---- flat parameter record consumed by the scoring kernel --------------
    def device_params(self) -> dict:
        sp = self.scaling_params
        w = self.penalty_weight
        clip = self.clip_range
        return {"kind": self.kind,
                "penalty_weight": math.nan if w is None else float(w),
                "clip_lo": math.nan if not clip else float(clip[0]),
                "clip_hi": math.nan if not clip else float(clip[1]),
                "objective_factor": float(sp["objective_factor"]),
                "objective_bias": float(sp["objective_bias"]),
                "penalty_factor": float(sp["penalty_factor"]),
                "penalty_bias": float(sp["penalty_bias"]),
                "valid_reward": float(getattr(self, "valid_reward", 0.0)),
                "invalid_penalty": float(getattr(self, "invalid_penalty", 0.0)),
                "invalid_objective_share": float(getattr(self, "invalid_objective_share", 1.0))}


class Summation(RewardFunction):
    kind = SUMMATION


class Replacement(RewardFunction):
    kind = REPLACEMENT

    def __init__(self, valid_reward: float = 1.0, **kw):
        super().__init__(**kw)
        if isinstance(valid_reward, str):
            # reference :237-239 calls an undefined helper (A.6 quirk 2)
            raise NotImplementedError("heuristic valid_reward is broken in the reference; pass a number")
        self.valid_reward = valid_reward

    def adjust_objective(self, objective, valid):
        return objective + self.valid_reward if valid else 0.0


class Parameterized(RewardFunction):
    kind = PARAMETERIZED

    def __init__(self, valid_reward: float = 0.0, invalid_penalty: float = 0.5,
                 invalid_objective_share: float = 1.0, **kw):
        super().__init__(**kw)
        for v in (valid_reward, invalid_penalty):
            if isinstance(v, str):
                raise NotImplementedError("heuristic offsets are broken in the reference (A.6 quirk 2)")
        assert valid_reward >= 0, "Valid reward must be >= 0"
        assert invalid_penalty >= 0, "Invalid penalty must be >= 0"
        assert 0 <= invalid_objective_share <= 1, "Objective share must be in [0, 1]"
        self.valid_reward = valid_reward
        self.invalid_penalty = invalid_penalty
        self.invalid_objective_share = invalid_objective_share

    def adjust_penalty(self, penalty, valid):
        return penalty + self.valid_reward if valid else penalty - self.invalid_penalty

    def adjust_objective(self, objective, valid):
        return objective if valid else objective * self.invalid_objective_share

    def calculate_cost(self, penalty, valid):
        return 0.0 if valid else super().calculate_cost(penalty, valid) + self.invalid_penalty


class OnlyObjective(RewardFunction):
    kind = ONLY_OBJECTIVE

    def __init__(self, **kw):
        super().__init__(penalty_weight=0.0, **kw)

    def adjust_penalty(self, penalty, valid):
        return 0.0


def load_reward_class(name: str):
    """String lookup with ``.capitalize()`` fallback, as reference
    ``opfgym/util/import_class.py:6-17`` does for ``reward_function='summation'``."""
    g = globals()
    for cand in (name, name.capitalize()):
        cls = g.get(cand)
        if isinstance(cls, type) and issubclass(cls, RewardFunction):
            return cls
    raise AttributeError(f"Class {name} not found in module opfgym_b200.reward!")
