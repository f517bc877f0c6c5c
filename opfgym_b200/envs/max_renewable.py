"""MaxRenewable on the batched engine (reference ``opfgym/envs/max_renewable.py:8-105``):
maximise renewable feed-in; actions = active power of the larger sgens/storages."""
from __future__ import annotations

from .. import net as pn
from ..grids import build_simbench_net
from ..opf_env import BatchedOpfEnv, split_build_kwargs


class MaxRenewable(BatchedOpfEnv):
    def __init__(self, simbench_network_name="1-HV-mixed--1-sw", gen_scaling=0.8,
                 load_scaling=0.8, min_storage_power=10, min_sgen_power=24, num_envs=1, **kwargs):
        self.min_sgen_power = min_sgen_power
        self.min_storage_power = min_storage_power
        build_kw = split_build_kwargs(kwargs)
        net, profiles = self._define_opf(simbench_network_name, gen_scaling=gen_scaling,
                                         load_scaling=load_scaling, **build_kw)
        free_storage = net.storage.index[~net.storage.controllable]
        obs_keys = [("sgen", "max_p_mw", net.sgen.index), ("load", "p_mw", net.load.index),
                    ("load", "q_mvar", net.load.index), ("storage", "p_mw", free_storage)]
        state_keys = [("sgen", "p_mw", net.sgen.index), ("load", "p_mw", net.load.index),
                      ("load", "q_mvar", net.load.index), ("storage", "p_mw", free_storage)]
        act_keys = [("sgen", "p_mw", net.sgen.index[net.sgen.controllable]),
                    ("storage", "p_mw", net.storage.index[net.storage.controllable])]
        super().__init__(net, act_keys, obs_keys, state_keys=state_keys, profiles=profiles,
                         num_envs=num_envs, **kwargs)

    def _define_opf(self, simbench_network_name, **kwargs):
        net, profiles = build_simbench_net(simbench_network_name, **kwargs)
        if len(net.ext_grid) > 1:
            net.ext_grid = net.ext_grid.iloc[0:1]
        net.trafo["max_loading_percent"] = 100.0
        net.load["controllable"] = False
        net.ext_grid["vm_pu"] = 1.0
        net.storage["controllable"] = net.storage.max_max_p_mw > self.min_storage_power
        net.storage["max_p_mw"] = net.storage["max_max_p_mw"]
        net.storage["min_p_mw"] = net.storage["min_min_p_mw"]
        net.sgen["controllable"] = net.sgen.max_max_p_mw > self.min_sgen_power
        net.sgen["min_p_mw"] = 0.0
        for unit in ("storage", "sgen"):
            net[unit]["q_mvar"] = 0.0
            net[unit]["max_q_mvar"] = 0.0
            net[unit]["min_q_mvar"] = 0.0
        for idx in net.sgen.index:
            pn.create_poly_cost(net, idx, "sgen", cp1_eur_per_mw=-30 / 1000)
        return net, profiles

    def _dynamic_columns(self):
        return [("sgen", "max_p_mw")]

    def _sampling(self, *args, **kwargs):
        super()._sampling(*args, **kwargs)
        self.run_row_program("mr_bounds", "sgen", lambda r: r.store(
            "max_p_mw", r.col("p_mw") * r.col("scaling") + 1e-6))
