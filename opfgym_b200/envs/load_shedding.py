"""LoadShedding on the batched engine (reference ``opfgym/envs/load_shedding.py:16-149``):
active power of the larger loads and storages is the action; shedding prices and
storage prices are sampled per episode; storage costs are piece-wise linear with
a charging/discharging efficiency."""
from __future__ import annotations

import numpy as np

from .. import net as pn
from ..grids import build_simbench_net
from ..opf_env import BatchedOpfEnv, split_build_kwargs


class LoadShedding(BatchedOpfEnv):
    def __init__(self, simbench_network_name="1-MV-comm--2-sw", gen_scaling=1.6,
                 load_scaling=2.2, min_load_power=0.6, min_storage_power=1.0,
                 max_p_exchange=8.0, storage_efficiency=0.95, num_envs=1, **kwargs):
        self.min_load_power = min_load_power
        self.min_storage_power = min_storage_power
        self.max_p_exchange = max_p_exchange
        self.storage_efficiency = storage_efficiency
        build_kw = split_build_kwargs(kwargs)
        net, profiles = self._define_opf(simbench_network_name, gen_scaling=gen_scaling,
                                         load_scaling=load_scaling, **build_kw)
        free_storage = net.storage.index[~net.storage.controllable]
        obs_keys = [("sgen", "p_mw", net.sgen.index), ("load", "max_p_mw", net.load.index),
                    ("load", "q_mvar", net.load.index), ("storage", "p_mw", free_storage),
                    ("poly_cost", "cp1_eur_per_mw", net.poly_cost.index),
                    ("pwl_cost", "cp1_eur_per_mw", net.pwl_cost.index)]
        state_keys = [("sgen", "p_mw", net.sgen.index), ("load", "p_mw", net.load.index),
                      ("load", "q_mvar", net.load.index), ("storage", "p_mw", free_storage)]
        act_keys = [("load", "p_mw", net.load.index[net.load.controllable]),
                    ("storage", "p_mw", net.storage.index[net.storage.controllable])]
        act_keys += self._extra_action_keys(net)
        super().__init__(net, act_keys, obs_keys, state_keys=state_keys, profiles=profiles,
                         num_envs=num_envs, pwl_price_columns=["price_charge", "price_discharge"],
                         **kwargs)

    def _extra_action_keys(self, net):
        return []

    def _define_opf(self, simbench_network_name, **kwargs):
        net, profiles = build_simbench_net(simbench_network_name, **kwargs)
        net.load["controllable"] = net.load.max_max_p_mw > self.min_load_power
        net.load["min_min_p_mw"] = 0.0          # every load can be shed completely
        net.load["min_p_mw"] = 0.0
        top = np.maximum(net.storage.min_min_p_mw.abs(), net.storage.max_max_p_mw.abs())
        for col, sign in (("min_p_mw", -1), ("max_p_mw", 1), ("min_min_p_mw", -1), ("max_max_p_mw", 1)):
            net.storage[col] = sign * top
        net.storage["controllable"] = net.storage.max_max_p_mw > self.min_storage_power
        net.sgen["controllable"] = False
        net.ext_grid["max_p_mw"] = self.max_p_exchange
        net.ext_grid["min_p_mw"] = -np.inf
        for idx in net.load.index[net.load.controllable]:
            pn.create_poly_cost(net, idx, "load", cp1_eur_per_mw=0)
        for idx in net.storage.index[net.storage.controllable]:
            pn.create_pwl_cost(net, idx, "storage", points=[[-1000, 0, 1], [0, 1000, 1]])
        net.poly_cost["min_cp1_eur_per_mw"] = -10.0   # shedding price range
        net.poly_cost["max_cp1_eur_per_mw"] = 0.0
        net.pwl_cost["cp1_eur_per_mw"] = 0.0
        net.pwl_cost["min_cp1_eur_per_mw"] = 0.0
        net.pwl_cost["max_cp1_eur_per_mw"] = 2.0      # storage price range
        net.pwl_cost["price_charge"] = 1.0
        net.pwl_cost["price_discharge"] = 1.0
        net.ext_grid["vm_pu"] = 1.0
        return net, profiles

    def _dynamic_columns(self):
        cols = [("poly_cost", "cp1_eur_per_mw"), ("pwl_cost", "cp1_eur_per_mw"),
                ("pwl_cost", "price_charge"), ("pwl_cost", "price_discharge"),
                ("load", "max_p_mw")]
        cols += [(u, c) for u in ("load", "storage") for c in ("max_q_mvar", "min_q_mvar")]
        return cols

    def _sampling(self, *args, **kwargs):
        super()._sampling(*args, **kwargs)
        self._sample_from_range("poly_cost", "cp1_eur_per_mw", self.net.poly_cost.index)
        self._sample_from_range("pwl_cost", "cp1_eur_per_mw", self.net.pwl_cost.index)
        eta = self.storage_efficiency

        def prices(r):      # pwl points [[-1000, 0, price*eta], [0, 1000, price/eta]]
            r.store("price_charge", r.col("cp1_eur_per_mw") * eta)
            r.store("price_discharge", r.col("cp1_eur_per_mw") / eta)

        def load_bounds(r):
            r.store("max_p_mw", r.col("p_mw") * r.col("scaling") + 1e-9)
            q = r.col("q_mvar") * r.col("scaling")
            r.store("max_q_mvar", q + 1e-9)
            r.store("min_q_mvar", q - 1e-9)

        def storage_bounds(r):
            q = r.col("q_mvar") * r.col("scaling")
            r.store("max_q_mvar", q + 1e-9)
            r.store("min_q_mvar", q - 1e-9)

        self.run_row_program("ls_prices", "pwl_cost", prices)
        self.run_row_program("ls_load", "load", load_bounds)
        self.run_row_program("ls_storage", "storage", storage_bounds)


class LoadSheddingReconfiguration(LoadShedding):
    """BASELINE.json config 5: LoadShedding whose agent also moves the transformer tap changers and
    switches the normally-open tie lines -- per-environment Ybus values.  The reference LoadShedding
    has no such actuators (SURVEY.md 8d calls this an extension); their semantics are those of the
    reference's ``examples/network_reconfiguration.py:34-35, 49-60``: ``trafo.tap_pos`` and
    ``line.in_service`` columns with ``min_`` / ``max_`` bounds, rounded by ``_apply_actions``
    (``opf_env.py:476-481``)."""

    def __init__(self, *args, tap_range: int = 3, **kwargs):
        self.tap_range = int(tap_range)
        super().__init__(*args, **kwargs)

    def _define_opf(self, simbench_network_name, **kwargs):
        net, profiles = LoadShedding._define_opf(self, simbench_network_name, **kwargs)
        net.trafo["min_tap_pos"] = -float(self.tap_range)
        net.trafo["max_tap_pos"] = float(self.tap_range)
        net.line["min_in_service"] = 0.0
        net.line["max_in_service"] = 1.0
        return net, profiles

    def _extra_action_keys(self, net):
        ties = net.line.index[~net.line.in_service]
        return [("trafo", "tap_pos", net.trafo.index), ("line", "in_service", ties)]
