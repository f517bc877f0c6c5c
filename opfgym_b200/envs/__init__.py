"""The five benchmark environments (reference ``opfgym/envs/__init__.py:12-35``).
The reference registers ``*-v0`` ids with gymnasium; the same ids are registered
here when gymnasium is importable, and ``make(id, num_envs=...)`` works either way."""
from .eco_dispatch import EcoDispatch
from .load_shedding import LoadShedding, LoadSheddingReconfiguration
from .max_renewable import MaxRenewable
from .q_market import QMarket
from .voltage_control import VoltageControl

REGISTRY = {"MaxRenewable-v0": MaxRenewable, "QMarket-v0": QMarket,
            "VoltageControl-v0": VoltageControl, "EcoDispatch-v0": EcoDispatch,
            "LoadShedding-v0": LoadShedding}


def make(env_id: str, num_envs: int = 1, **kwargs):
    return REGISTRY[env_id](num_envs=num_envs, **kwargs)


try:  # pragma: no cover - gymnasium is absent from the build image
    from gymnasium.envs.registration import register
    for _id, _cls in REGISTRY.items():
        register(id=f"B200-{_id}", entry_point=f"opfgym_b200.envs:{_cls.__name__}")
except Exception:  # noqa: BLE001
    pass
