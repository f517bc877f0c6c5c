"""opfgym_b200 -- B200-native batched AC power-flow + reward engine behind opfgym's OpfEnv API."""
__version__ = "0.1.0"

from .constraints import Constraint  # noqa: F401
from .reward import RewardFunction  # noqa: F401


def __getattr__(name):   # lazy: importing the package must not need torch
    if name in ("BatchedOpfEnv", "PowerFlowNotAvailable"):
        from . import opf_env
        return getattr(opf_env, name)
    raise AttributeError(name)
