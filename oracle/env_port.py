"""CPU ORACLE -- TEST / BASELINE INFRASTRUCTURE ONLY.

Single-environment CPU port of the reference hot path ``OpfEnv.reset`` +
``OpfEnv.step`` (``opfgym/opf_env.py:177-220, 374-419``) for the VoltageControl
benchmark (``opfgym/envs/voltage_control.py``): one pandas-backed net per
environment, uniform sampling, the env's ``_sampling`` hook, ``_apply_actions``,
the oracle power flow (``oracle/pf.py`` standing in for ``pp.runpp``), then the
oracle scoring and the observation gather.  ``bench.py`` times it on the host
cores as the ``cpu_baseline`` / ``--impl reference`` arm (the real pandapower
path is not installable; this port omits pandapower's pandas<->ppc overhead and
therefore OVER-states the CPU path's speed).
"""
from __future__ import annotations

import numpy as np

from opfgym_b200 import constraints as C
from opfgym_b200 import reward as R
from opfgym_b200.net import LoadflowNotConverged
from opfgym_b200.ppc import PpcBuilder
from oracle import pf, scoring


class _Params:
    min_sgen_power = 0.5
    min_storage_power = 0.5
    cos_phi = 0.95
    market_based = False
    max_q_exchange = 0.5


class OracleVoltageControl:
    def __init__(self, seed=0, n_profile_steps=672):
        from opfgym_b200.envs.voltage_control import VoltageControl
        self.net, _ = VoltageControl._define_opf(_Params(), "1-MV-semiurb--1-sw", gen_scaling=1.3,
                                                 load_scaling=1.5, n_profile_steps=n_profile_steps)
        net = self.net
        self.obs_keys = [("sgen", "p_mw", net.sgen.index), ("storage", "p_mw", net.storage.index),
                         ("load", "p_mw", net.load.index), ("load", "q_mvar", net.load.index)]
        self.act_keys = [("sgen", "q_mvar", net.sgen.index[net.sgen.controllable]),
                         ("storage", "q_mvar", net.storage.index[net.storage.controllable])]
        self.n_act = sum(len(i) for _, _, i in self.act_keys)
        self.constraints = C.create_default_constraints(net, {})
        self.reward_function = R.Summation()
        self.builder = PpcBuilder(net)
        self.rng = np.random.default_rng(seed)

    def reset(self):
        net = self.net
        for table, column, idxs in self.obs_keys:               # _sample_uniform :253-284
            df = net[table]
            if not len(idxs):
                continue
            r = self.rng.uniform(df[f"min_min_{column}"].loc[idxs], df[f"max_max_{column}"].loc[idxs])
            net[table].loc[idxs, column] = r / df.scaling.loc[idxs]
        for unit in ("sgen", "storage"):                        # voltage_control.py:121-133
            df = net[unit]
            df["max_p_mw"] = df.p_mw * df.scaling + 1e-9
            df["min_p_mw"] = df.p_mw * df.scaling - 1e-9
            q_max = (df.max_s_mva ** 2 - df.max_p_mw ** 2) ** 0.5
            df["min_q_mvar"] = -q_max
            df["max_q_mvar"] = q_max
            df["q_mvar"] = 0.0
        self._apply_actions(np.full(self.n_act, 0.5))
        return self._obs()

    def _apply_actions(self, action):
        action = np.clip(action, 0.0, 1.0)
        k = 0
        for table, column, idxs in self.act_keys:               # :432-483
            df = self.net[table]
            a = action[k:k + len(idxs)]
            lo, hi = df[f"min_{column}"].loc[idxs], df[f"max_{column}"].loc[idxs]
            self.net[table].loc[idxs, column] = (a * (hi - lo).values + lo) / df.scaling.loc[idxs]
            k += len(idxs)

    def _obs(self):
        return np.concatenate([self.net[t].loc[i, c].to_numpy() for t, c, i in self.obs_keys])

    def step(self, action):
        self._apply_actions(action)
        try:
            pf.runpp(self.net, self.builder)
        except LoadflowNotConverged:
            return np.array([np.nan]), np.nan, True, False, {}
        out = scoring.step_reward(self.net, self.constraints, self.reward_function)
        info = {k: out[k] for k in ("valids", "violations", "unscaled_penalties", "cost")}
        return self._obs(), out["reward"], True, False, info


_ENV = None


def _worker_steps(args):
    """Run ``n`` (reset, step) pairs in this process; returns (n, n_converged)."""
    global _ENV
    seed, n = args
    if _ENV is None:
        _ENV = OracleVoltageControl(seed=seed)
    env = _ENV
    ok = 0
    for _ in range(n):
        env.reset()
        _, reward, _, _, _ = env.step(env.rng.uniform(0, 1, env.n_act))
        ok += int(reward == reward)
    return n, ok
