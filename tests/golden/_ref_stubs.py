"""Stub modules that let the UNMODIFIED reference package (/root/reference/opfgym)
be imported in the build container, where pandapower / simbench / gymnasium are
not installable.  Used ONLY by ``make_golden.py`` to generate the fixtures in
this directory; nothing in the product or in the GPU tests imports it.

What is stubbed, and with what:
* ``pandapower``  -> the Net container of opfgym_b200.net (table shape only),
                     ``runpp`` = the CPU oracle (oracle/pf.py).  Every other
                     line that runs is the reference's own code.
* ``simbench``    -> returns the synthetic stand-in grid / profiles of the same name.
* ``gymnasium``   -> ``Env``, ``spaces.Box``, ``ObservationWrapper``, ``register``
                     with gymnasium's documented semantics (float32 bounds,
                     ``np_random`` = numpy Generator seeded in ``reset``).
"""
import sys
import types

import numpy as np


def install(profile_steps=None):
    from opfgym_b200 import grids
    from opfgym_b200 import net as pn
    from oracle import pf

    # ------------------------------------------------------------ pandapower
    pp = types.ModuleType("pandapower")
    pp.pandapowerNet = pn.Net
    pp.powerflow = types.ModuleType("pandapower.powerflow")
    pp.powerflow.LoadflowNotConverged = pn.LoadflowNotConverged
    pp.optimal_powerflow = types.ModuleType("pandapower.optimal_powerflow")

    class OPFNotConverged(Exception):
        pass
    pp.optimal_powerflow.OPFNotConverged = OPFNotConverged
    pp.create_poly_cost = pn.create_poly_cost
    pp.create_pwl_cost = pn.create_pwl_cost
    pp.diagnostic = lambda net: "diagnostic unavailable (stub)"

    def runpp(net, enforce_q_lims=False, lightsim2grid=None, **kw):
        return pf.runpp(net, enforce_q_lims=enforce_q_lims, **kw)
    pp.runpp = runpp

    def runopp(net, **kw):
        raise OPFNotConverged("no OPF in the stub")
    pp.runopp = runopp
    pp.networks = types.ModuleType("pandapower.networks")
    pp.plotting = types.ModuleType("pandapower.plotting")
    for name, mod in (("pandapower", pp), ("pandapower.powerflow", pp.powerflow),
                      ("pandapower.optimal_powerflow", pp.optimal_powerflow),
                      ("pandapower.networks", pp.networks),
                      ("pandapower.plotting", pp.plotting)):
        sys.modules[name] = mod

    # -------------------------------------------------------------- simbench
    sb = types.ModuleType("simbench")
    sb.get_simbench_net = lambda name: grids.raw_standin(name)
    sb.profiles_are_missing = lambda net: False

    def get_absolute_values(net, profiles_instead_of_study_cases=True):
        return grids.synth_profiles(net, n_steps=profile_steps or grids.N_SIMBENCH_STEPS)
    sb.get_absolute_values = get_absolute_values
    sys.modules["simbench"] = sb

    # ------------------------------------------------------------- gymnasium
    gym = types.ModuleType("gymnasium")

    class Box:
        def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
            if shape is None:
                shape = np.broadcast(np.asarray(low), np.asarray(high)).shape
            self.shape = tuple(shape)
            self.dtype = np.dtype(dtype)
            self.low = np.broadcast_to(np.asarray(low, dtype=float), self.shape).astype(self.dtype)
            self.high = np.broadcast_to(np.asarray(high, dtype=float), self.shape).astype(self.dtype)
            self._rng = np.random.default_rng(seed)

        def sample(self):
            return self._rng.uniform(self.low, self.high).astype(self.dtype)

        def seed(self, seed=None):
            self._rng = np.random.default_rng(seed)

        def contains(self, x):
            return bool(np.all(x >= self.low) and np.all(x <= self.high))

    class Env:
        _np_random = None

        @property
        def np_random(self):
            if self._np_random is None:
                self._np_random = np.random.default_rng()
            return self._np_random

        @np_random.setter
        def np_random(self, value):
            self._np_random = value

        def reset(self, *, seed=None, options=None):
            if seed is not None:
                self._np_random = np.random.default_rng(seed)

    class Wrapper(Env):
        def __init__(self, env):
            self.env = env
            self.observation_space = env.observation_space
            self.action_space = env.action_space

        def __getattr__(self, name):
            return getattr(self.env, name)

    class ObservationWrapper(Wrapper):
        def reset(self, *, seed=None, options=None):
            Env.reset(self, seed=seed)
            obs, info = self.env.reset(seed=seed, options=options)
            return self.observation(obs), info

        def step(self, action):
            obs, r, term, trunc, info = self.env.step(action)
            return self.observation(obs), r, term, trunc, info

    gym.Env = Env
    gym.Wrapper = Wrapper
    gym.ObservationWrapper = ObservationWrapper
    gym.spaces = types.ModuleType("gymnasium.spaces")
    gym.spaces.Box = Box
    gym.envs = types.ModuleType("gymnasium.envs")
    gym.envs.registration = types.ModuleType("gymnasium.envs.registration")
    gym.envs.registration.register = lambda **kw: None
    for name, mod in (("gymnasium", gym), ("gymnasium.spaces", gym.spaces),
                      ("gymnasium.envs", gym.envs),
                      ("gymnasium.envs.registration", gym.envs.registration)):
        sys.modules[name] = mod

    if "/root/reference" not in sys.path:
        sys.path.append("/root/reference")
