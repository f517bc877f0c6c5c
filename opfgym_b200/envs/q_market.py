"""QMarket = VoltageControl with sampled reactive-power prices (reference
``opfgym/envs/q_market.py:5-36``)."""
from .voltage_control import VoltageControl


class QMarket(VoltageControl):
    def __init__(self, simbench_network_name="1-MV-rural--0-sw", gen_scaling=1.0,
                 load_scaling=1.5, min_sgen_power=0.2, cos_phi=0.95, max_q_exchange=0.1,
                 market_based=True, **kwargs):
        super().__init__(simbench_network_name=simbench_network_name, load_scaling=load_scaling,
                         gen_scaling=gen_scaling, cos_phi=cos_phi, max_q_exchange=max_q_exchange,
                         market_based=market_based, min_sgen_power=min_sgen_power, **kwargs)
