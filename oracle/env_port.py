"""CPU ORACLE -- TEST / BASELINE INFRASTRUCTURE ONLY.

Single-environment CPU port of the reference hot path ``OpfEnv.reset`` +
``OpfEnv.step`` (``opfgym/opf_env.py:177-220, 374-419``) for the benchmark
environments (``OracleVoltageControl`` for the headline config, ``OracleEnv`` for
all five + the config-5 variant): one pandas-backed net per
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


class _Obj:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class OracleEnv:
    """The same single-environment loop for any of the benchmark environments (``kind`` =
    VoltageControl, QMarket, EcoDispatch, MaxRenewable, LoadShedding, LoadSheddingReconfiguration):
    ``reset`` = uniform sampling of the state keys + the environment's ``_sampling`` hook restated
    from the reference (``envs/voltage_control.py:111-133``, ``eco_dispatch.py:111-123``,
    ``load_shedding.py:122-149``, ``max_renewable.py:101-105``) + centre action; ``step`` = actions,
    oracle power flow, oracle scoring, observation."""

    def __init__(self, kind="VoltageControl", seed=0, n_profile_steps=672):
        from opfgym_b200 import envs as E
        self.kind = kind
        cls = getattr(E, kind)
        import inspect
        defaults = {k: v.default for k, v in inspect.signature(cls.__init__).parameters.items()
                    if v.default is not inspect.Parameter.empty}
        if kind == "QMarket":      # QMarket only overrides VoltageControl's defaults
            defaults = {**{k: v.default for k, v in inspect.signature(E.VoltageControl.__init__).parameters.items()
                           if v.default is not inspect.Parameter.empty}, **defaults}
        if kind == "LoadSheddingReconfiguration":
            defaults = {**{k: v.default for k, v in inspect.signature(E.LoadShedding.__init__).parameters.items()
                           if v.default is not inspect.Parameter.empty}, **defaults}
        self.p = p = _Obj(**defaults)
        net, _ = cls._define_opf(p, p.simbench_network_name, gen_scaling=p.gen_scaling,
                                 load_scaling=p.load_scaling, n_profile_steps=n_profile_steps)
        self.net = net
        ctrl = lambda t: net[t].index[net[t].controllable.astype(bool)]
        free = lambda t: net[t].index[~net[t].controllable.astype(bool)]
        if kind in ("VoltageControl", "QMarket"):
            self.state_keys = [("sgen", "p_mw", net.sgen.index), ("storage", "p_mw", net.storage.index),
                               ("load", "p_mw", net.load.index), ("load", "q_mvar", net.load.index)]
            self.obs_keys = list(self.state_keys)
            if p.market_based:
                self.obs_keys.append(("poly_cost", "cq2_eur_per_mvar2", net.poly_cost.index))
            self.act_keys = [("sgen", "q_mvar", ctrl("sgen")), ("storage", "q_mvar", ctrl("storage"))]
        elif kind == "EcoDispatch":
            self.obs_keys = [("load", "p_mw", net.load.index), ("load", "q_mvar", net.load.index),
                             ("poly_cost", "cp1_eur_per_mw", net.poly_cost.index),
                             ("pwl_cost", "cp1_eur_per_mw", net.pwl_cost.index),
                             ("sgen", "p_mw", free("sgen")), ("storage", "p_mw", net.storage.index),
                             ("storage", "q_mvar", net.storage.index)]
            self.state_keys = [k for k in self.obs_keys if "cost" not in k[0]]
            self.act_keys = [("sgen", "p_mw", ctrl("sgen")), ("gen", "p_mw", ctrl("gen"))]
        elif kind == "MaxRenewable":
            self.state_keys = [("sgen", "p_mw", net.sgen.index), ("load", "p_mw", net.load.index),
                               ("load", "q_mvar", net.load.index), ("storage", "p_mw", free("storage"))]
            self.obs_keys = [("sgen", "max_p_mw", net.sgen.index)] + self.state_keys[1:]
            self.act_keys = [("sgen", "p_mw", ctrl("sgen")), ("storage", "p_mw", ctrl("storage"))]
        else:                      # LoadShedding (+ reconfiguration actuators)
            self.state_keys = [("sgen", "p_mw", net.sgen.index), ("load", "p_mw", net.load.index),
                               ("load", "q_mvar", net.load.index), ("storage", "p_mw", free("storage"))]
            self.obs_keys = [("sgen", "p_mw", net.sgen.index), ("load", "max_p_mw", net.load.index),
                             ("load", "q_mvar", net.load.index), ("storage", "p_mw", free("storage")),
                             ("poly_cost", "cp1_eur_per_mw", net.poly_cost.index),
                             ("pwl_cost", "cp1_eur_per_mw", net.pwl_cost.index)]
            self.act_keys = [("load", "p_mw", ctrl("load")), ("storage", "p_mw", ctrl("storage"))]
            if kind == "LoadSheddingReconfiguration":
                self.act_keys += [("trafo", "tap_pos", net.trafo.index),
                                  ("line", "in_service", net.line.index[~net.line.in_service.astype(bool)])]
        self.act_keys = [k for k in self.act_keys if len(k[2])]
        self.n_act = sum(len(i) for _, _, i in self.act_keys)
        self.constraints = C.create_default_constraints(net, {})
        self.reward_function = R.Summation()
        self.dynamic_topology = kind == "LoadSheddingReconfiguration"
        self.builder = None if self.dynamic_topology else PpcBuilder(net)
        self.rng = np.random.default_rng(seed)

    def _sample(self, table, column, idxs):
        df = self.net[table]
        if not len(idxs):
            return
        lo = df[f"min_min_{column}"] if f"min_min_{column}" in df.columns else df[f"min_{column}"]
        hi = df[f"max_max_{column}"] if f"max_max_{column}" in df.columns else df[f"max_{column}"]
        r = self.rng.uniform(lo.loc[idxs].to_numpy(float), hi.loc[idxs].to_numpy(float))
        scal = df.scaling.loc[idxs].to_numpy(float) if "scaling" in df.columns else 1.0
        self.net[table].loc[idxs, column] = r / scal

    def reset(self):
        net, kind = self.net, self.kind
        for table, column, idxs in self.state_keys:             # _sample_uniform, opf_env.py:253-284
            self._sample(table, column, idxs)
        if kind in ("VoltageControl", "QMarket"):               # voltage_control.py:111-133
            if self.p.market_based:
                self._sample("poly_cost", "cq2_eur_per_mvar2", net.poly_cost.index)
            for unit in ("sgen", "storage"):
                df = net[unit]
                df["max_p_mw"] = df.p_mw * df.scaling + 1e-9
                df["min_p_mw"] = df.p_mw * df.scaling - 1e-9
                q_max = (df.max_s_mva ** 2 - df.max_p_mw ** 2) ** 0.5
                df["min_q_mvar"] = -q_max
                df["max_q_mvar"] = q_max
                df["q_mvar"] = 0.0
        elif kind == "EcoDispatch":                             # eco_dispatch.py:111-123
            self._sample("poly_cost", "cp1_eur_per_mw", net.poly_cost.index)
            self._sample("pwl_cost", "cp1_eur_per_mw", net.pwl_cost.index)
            for idx in net.pwl_cost.index:
                net.pwl_cost.at[idx, "points"] = [[0, 10000, net.pwl_cost.at[idx, "cp1_eur_per_mw"]]]
        elif kind == "MaxRenewable":                            # max_renewable.py:101-105
            net.sgen["max_p_mw"] = net.sgen.p_mw * net.sgen.scaling + 1e-6
        else:                                                   # load_shedding.py:122-149
            self._sample("poly_cost", "cp1_eur_per_mw", net.poly_cost.index)
            self._sample("pwl_cost", "cp1_eur_per_mw", net.pwl_cost.index)
            eta = self.p.storage_efficiency
            for idx in net.pwl_cost.index:
                price = net.pwl_cost.at[idx, "cp1_eur_per_mw"]
                net.pwl_cost.at[idx, "points"] = [[-1000, 0, price * eta], [0, 1000, price / eta]]
            net.load["max_p_mw"] = net.load.p_mw * net.load.scaling + 1e-9
            for unit in ("load", "storage"):
                df = net[unit]
                df["max_q_mvar"] = df.q_mvar * df.scaling + 1e-9
                df["min_q_mvar"] = df.q_mvar * df.scaling - 1e-9
        self._apply_actions(np.full(self.n_act, 0.5))
        return self._obs()

    def _apply_actions(self, action):
        action = np.clip(action, 0.0, 1.0)                      # opf_env.py:429
        k = 0
        for table, column, idxs in self.act_keys:               # :432-483
            df = self.net[table]
            a = action[k:k + len(idxs)]
            lo, hi = df[f"min_{column}"].loc[idxs].to_numpy(float), df[f"max_{column}"].loc[idxs].to_numpy(float)
            sp = a * (hi - lo) + lo
            if "scaling" in df.columns:
                sp = sp / df.scaling.loc[idxs].to_numpy(float)
            if column in ("closed", "in_service"):
                sp = np.round(sp).astype(bool)
            elif column in ("tap_pos", "step"):
                sp = np.round(sp)
            self.net[table].loc[idxs, column] = sp
            k += len(idxs)

    def _obs(self):
        return np.concatenate([self.net[t].loc[i, c].to_numpy(float) for t, c, i in self.obs_keys])

    def step(self, action):
        self._apply_actions(action)
        try:
            pf.runpp(self.net, PpcBuilder(self.net) if self.dynamic_topology else self.builder)
        except LoadflowNotConverged:
            return np.array([np.nan]), np.nan, True, False, {}
        out = scoring.step_reward(self.net, self.constraints, self.reward_function)
        info = {k: out[k] for k in ("valids", "violations", "unscaled_penalties", "cost")}
        return self._obs(), out["reward"], True, False, info


_ENV = {}


def _worker_steps(args):
    """Run ``n`` (reset, step) pairs in this process; returns (n, n_converged)."""
    seed, n = args[0], args[1]
    kind = args[2] if len(args) > 2 else "VoltageControl"
    if kind not in _ENV:
        _ENV[kind] = OracleVoltageControl(seed=seed) if kind == "VoltageControl" else OracleEnv(kind, seed=seed)
    env = _ENV[kind]
    ok = 0
    for _ in range(n):
        env.reset()
        _, reward, _, _, _ = env.step(env.rng.uniform(0, 1, env.n_act))
        ok += int(reward == reward)
    return n, ok
