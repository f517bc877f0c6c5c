"""opfgym_b200 -- B200-native batched AC power-flow + reward engine behind opfgym's OpfEnv API."""
__version__ = "0.1.0"
