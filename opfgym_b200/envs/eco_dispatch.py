"""EcoDispatch on the batched engine (reference ``opfgym/envs/eco_dispatch.py:8-123``):
active-power set-points of all generators are the actions, prices are sampled per
episode, the slack is priced through a one-segment piece-wise-linear cost."""
from __future__ import annotations

from .. import net as pn
from ..grids import build_simbench_net
from ..opf_env import BatchedOpfEnv, split_build_kwargs


class EcoDispatch(BatchedOpfEnv):
    def __init__(self, simbench_network_name="1-HV-urban--0-sw", gen_scaling=1.0,
                 load_scaling=1.5, max_price_eur_gwh=0.5, min_power=0, num_envs=1, **kwargs):
        self.max_price_eur_gwh = max_price_eur_gwh
        self.min_power = min_power
        build_kw = split_build_kwargs(kwargs)
        net, profiles = self._define_opf(simbench_network_name, gen_scaling=gen_scaling,
                                         load_scaling=load_scaling, **build_kw)
        obs_keys = [("load", "p_mw", net.load.index), ("load", "q_mvar", net.load.index),
                    ("poly_cost", "cp1_eur_per_mw", net.poly_cost.index),
                    ("pwl_cost", "cp1_eur_per_mw", net.pwl_cost.index),
                    ("sgen", "p_mw", net.sgen.index[~net.sgen.controllable]),
                    ("storage", "p_mw", net.storage.index),
                    ("storage", "q_mvar", net.storage.index)]
        act_keys = [("sgen", "p_mw", net.sgen.index[net.sgen.controllable]),
                    ("gen", "p_mw", net.gen.index[net.gen.controllable])]
        super().__init__(net, act_keys, obs_keys, profiles=profiles, num_envs=num_envs,
                         pwl_price_columns=["cp1_eur_per_mw"], **kwargs)

    def _define_opf(self, simbench_network_name, **kwargs):
        net, profiles = build_simbench_net(simbench_network_name, **kwargs)
        net.ext_grid["vm_pu"] = 1.0
        net.gen["vm_pu"] = 1.0
        net.load["controllable"] = False
        net.ext_grid["min_p_mw"] = 0.0                              # no selling to the upper grid
        net.ext_grid["max_p_mw"] = float(net.sgen.max_max_p_mw.max())
        for unit in ("sgen", "gen"):
            net[unit]["min_p_mw"] = 0.0
            net[unit]["max_p_mw"] = net[unit]["max_max_p_mw"]
            net[unit]["max_q_mvar"] = 0.0                           # reactive power neglected
            net[unit]["min_q_mvar"] = 0.0
        net.sgen["controllable"] = net.sgen.max_max_p_mw > self.min_power
        net.sgen["min_min_p_mw"] = 0.0
        net.gen["controllable"] = True
        for idx in net.ext_grid.index:
            pn.create_pwl_cost(net, idx, "ext_grid", points=[[0, 10000, 1]])
        for idx in net.sgen.index[net.sgen.controllable]:
            pn.create_poly_cost(net, idx, "sgen", cp1_eur_per_mw=0)
        for idx in net.gen.index[net.gen.controllable]:
            pn.create_poly_cost(net, idx, "gen", cp1_eur_per_mw=0)
        for table in ("poly_cost", "pwl_cost"):
            net[table]["min_cp1_eur_per_mw"] = 0.0
            net[table]["max_cp1_eur_per_mw"] = self.max_price_eur_gwh
        net.pwl_cost["cp1_eur_per_mw"] = 0.0
        return net, profiles

    def _dynamic_columns(self):
        return [("poly_cost", "cp1_eur_per_mw"), ("pwl_cost", "cp1_eur_per_mw")]

    def _sampling(self, *args, **kwargs):
        super()._sampling(*args, **kwargs)
        # prices ~ U(0, max_price); the slack's single pwl segment [0, 10000] takes the
        # sampled pwl_cost.cp1_eur_per_mw as its price (pwl_price_columns)
        self._sample_from_range("poly_cost", "cp1_eur_per_mw", self.net.poly_cost.index)
        self._sample_from_range("pwl_cost", "cp1_eur_per_mw", self.net.pwl_cost.index)
