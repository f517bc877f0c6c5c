from .stochastic_obs import StochasticObservation  # noqa: F401
