"""VoltageControl on the batched engine.

Same problem definition as reference ``opfgym/envs/voltage_control.py:8-133``
(SURVEY.md App. A.1): reactive-power set-points of the larger sgens/storages are
the actions; every load P/Q and every sgen/storage P is observed; objective =
loss costs (+ quadratic Q prices if ``market_based``); constraints = voltage
band, line/trafo loading, slack reactive exchange.
"""
from __future__ import annotations

from .. import net as pn
from ..grids import build_simbench_net
from ..opf_env import BatchedOpfEnv, split_build_kwargs


class VoltageControl(BatchedOpfEnv):
    def __init__(self, simbench_network_name="1-MV-semiurb--1-sw", load_scaling=1.5,
                 gen_scaling=1.3, cos_phi=0.95, max_q_exchange=0.5, min_sgen_power=0.5,
                 min_storage_power=0.5, market_based=False, num_envs=1, **kwargs):
        self.min_sgen_power = min_sgen_power
        self.min_storage_power = min_storage_power
        self.cos_phi = cos_phi
        self.market_based = market_based
        self.max_q_exchange = max_q_exchange
        build_kw = split_build_kwargs(kwargs)
        net, profiles = self._define_opf(simbench_network_name, gen_scaling=gen_scaling,
                                         load_scaling=load_scaling, **build_kw)
        obs_keys = [("sgen", "p_mw", net.sgen.index), ("storage", "p_mw", net.storage.index),
                    ("load", "p_mw", net.load.index), ("load", "q_mvar", net.load.index)]
        if market_based:
            obs_keys.append(("poly_cost", "cq2_eur_per_mvar2", net.poly_cost.index))
        act_keys = [("sgen", "q_mvar", net.sgen.index[net.sgen.controllable]),
                    ("storage", "q_mvar", net.storage.index[net.storage.controllable])]
        super().__init__(net, act_keys, obs_keys, profiles=profiles, num_envs=num_envs, **kwargs)

    def _define_opf(self, simbench_network_name, **kwargs):
        net, profiles = build_simbench_net(simbench_network_name, **kwargs)
        net.load["controllable"] = False
        for unit, threshold in (("sgen", self.min_sgen_power), ("storage", self.min_storage_power)):
            df = net[unit]
            df["controllable"] = df.max_max_p_mw > threshold
            # sgens may exceed their active rating in Q (1/cos_phi); storages Q range = P range
            df["max_s_mva"] = df.max_max_p_mw / self.cos_phi if unit == "sgen" else df.max_max_p_mw.abs()
            df["max_max_q_mvar"] = df.max_s_mva
            df["min_min_q_mvar"] = -df.max_s_mva
        net.ext_grid["max_q_mvar"] = self.max_q_exchange
        net.ext_grid["min_q_mvar"] = -self.max_q_exchange
        self.loss_costs = 0.03   # eur/1000 per MW
        for unit, sign in (("sgen", 1.0), ("storage", -1.0)):
            for idx in net[unit].index[net[unit].controllable]:
                pn.create_poly_cost(net, idx, unit, cp1_eur_per_mw=sign * self.loss_costs,
                                    cq2_eur_per_mvar2=0)
        for idx in net.ext_grid.index:
            pn.create_poly_cost(net, idx, "ext_grid", cp1_eur_per_mw=self.loss_costs,
                                cq2_eur_per_mvar2=0)
        assert len(net.gen) == 0
        self.max_price = 0.03
        net.poly_cost["min_cq2_eur_per_mvar2"] = 0.0
        net.poly_cost["max_cq2_eur_per_mvar2"] = self.max_price
        return net, profiles

    def _dynamic_columns(self):
        cols = [(u, c) for u in ("sgen", "storage")
                for c in ("max_p_mw", "min_p_mw", "min_q_mvar", "max_q_mvar", "q_mvar")]
        if self.market_based:
            cols.append(("poly_cost", "cq2_eur_per_mvar2"))
        return cols

    def _sampling(self, *args, **kwargs):
        super()._sampling(*args, **kwargs)
        if self.market_based:   # reactive prices ~ U(0, max_price), drawn per element type
            pc = self.net.poly_cost
            for unit in ("sgen", "ext_grid", "storage"):
                self._sample_from_range("poly_cost", "cq2_eur_per_mvar2", pc.index[pc.et == unit])
        for unit in ("sgen", "storage"):
            self.run_row_program("vc_bounds", unit, self._bounds_program)

    @staticmethod
    def _bounds_program(r):
        # active power is not controllable: pin its bounds to the sampled value; offer the
        # whole remaining apparent-power capability as reactive range; start from Q = 0
        p_scaled = r.col("p_mw") * r.col("scaling")
        p_max = p_scaled + 1e-9
        r.store("max_p_mw", p_max)
        r.store("min_p_mw", p_scaled - 1e-9)
        q_max = (r.col("max_s_mva") ** 2 - p_max ** 2) ** 0.5
        r.store("min_q_mvar", -q_max)
        r.store("max_q_mvar", q_max)
        r.store("q_mvar", 0.0)
