"""Synthetic SimBench-like grids and profiles, plus the reference's net preparation.

SimBench data cannot be installed in the build image, so the
five benchmark grids (reference ``docs/source/benchmarks.rst:19-27``) are
replaced by deterministic synthetic stand-ins with the same bus counts and the
same observation/action dimensions: radial 20-kV feeders under two 110/20-kV
transformers with a 150 degree phase shift (``synth_mv``), and a meshed 110-kV
grid under 380/110-kV transformers (``synth_hv``).  Line/trafo parameters are
pandapower standard-type values.

``build_simbench_net`` mirrors reference ``opfgym/simbench/build_simbench_net.py:5-23``
(scaling columns, voltage band, max loading, profile repair and the
``min_min_/max_max_/mean_/std_dev_`` columns) on those stand-ins.
"""
from __future__ import annotations

import numpy as np
import pandas as pd

from . import net as pn

N_SIMBENCH_STEPS = 24 * 4 * 366  # reference opfgym/simbench/data_split.py:13

# r_ohm_per_km, x_ohm_per_km, c_nf_per_km, max_i_ka
_MV_CABLES = [(0.122, 0.112, 304.0, 0.421), (0.161, 0.117, 273.0, 0.362)]
_HV_LINES = [(0.1188, 0.39, 9.0, 0.645), (0.0949, 0.38, 9.2, 0.74)]

# simbench code -> synthetic stand-in spec
_STANDINS = {
    "1-MV-semiurb--1-sw": dict(kind="mv", nb=122, feeders=8, n_load=115, n_sgen=198,
                               n_storage=14, big_sgen=(10, 0.5, 1.3), big_storage=(4, 0.5, 1.0)),
    "1-MV-rural--0-sw": dict(kind="mv", nb=97, feeders=6, n_load=92, n_sgen=110,
                             n_storage=0, big_sgen=(10, 0.2, 1.0), big_storage=(0, 0.5, 1.0)),
    "1-MV-comm--2-sw": dict(kind="mv", nb=111, feeders=7, n_load=106, n_sgen=150,
                            n_storage=12, big_sgen=(6, 1.0, 1.6), big_storage=(4, 1.0, 1.0),
                            big_load=(12, 0.6, 2.2)),
    "1-HV-urban--0-sw": dict(kind="hv", nb=372, n_load=79, n_sgen=37, n_storage=0,
                             n_gen=5, big_sgen=(37, 0.0, 1.0), big_storage=(0, 10.0, 1.0)),
    "1-HV-mixed--1-sw": dict(kind="hv", nb=355, n_load=58, n_sgen=50, n_storage=8,
                             n_gen=0, big_sgen=(16, 24.0, 0.8), big_storage=(2, 10.0, 1.0)),
}


def standin_names():
    return tuple(_STANDINS)


# ----------------------------------------------------------------------------- grids
def synth_mv(nb=122, feeders=8, n_load=115, n_sgen=198, n_storage=14, n_ties=4,
             seed=0, name="synth_mv") -> pn.Net:
    """Radial 20-kV grid: bus 0 = 110 kV slack, buses 1/2 = busbars of two
    110/20-kV 40-MVA YNd5 transformers, the rest spread over tree-shaped feeders."""
    rng = np.random.default_rng(seed)
    net = pn.create_empty_network(name=name, sn_mva=1.0)
    pn.create_buses(net, 1, 110.0)
    pn.create_buses(net, nb - 1, 20.0)
    pn.create_ext_grid(net, 0, vm_pu=1.025)
    for lv in (1, 2):
        pn.create_transformer_from_parameters(
            net, 0, lv, sn_mva=40.0, vn_hv_kv=110.0, vn_lv_kv=20.0,
            vkr_percent=0.34, vk_percent=16.2, pfe_kw=18.0, i0_percent=0.05,
            shift_degree=150.0, tap_side="hv", tap_neutral=0.0, tap_min=-9.0,
            tap_max=9.0, tap_step_percent=1.5, tap_pos=0.0)
    n_feed = nb - 3
    sizes = np.full(feeders, n_feed // feeders)
    sizes[: n_feed % feeders] += 1
    fb, tb, ends = [], [], []
    nxt = 3
    for f, size in enumerate(sizes):
        members = []
        for k in range(size):
            if k == 0:
                parent = 1 + (f % 2)
            elif rng.random() < 0.8:
                parent = members[-1]
            else:
                parent = members[rng.integers(0, len(members))]
            fb.append(parent)
            tb.append(nxt)
            members.append(nxt)
            nxt += 1
        ends.append(members[-1])
    n_line = len(fb)
    typ = rng.integers(0, 2, n_line)
    par = np.array(_MV_CABLES)[typ]
    pn.create_lines_from_parameters(
        net, fb, tb, rng.uniform(0.5, 3.0, n_line), par[:, 0], par[:, 1], par[:, 2], par[:, 3])
    # normally-open ring ties between neighbouring feeder ends
    for k in range(min(n_ties, feeders - 1)):
        p = _MV_CABLES[0]
        pn.create_line_from_parameters(net, ends[k], ends[k + 1], 1.0, *p, in_service=False)
    mv_buses = np.arange(3, nb)
    _populate(net, rng, mv_buses, n_load, n_sgen, n_storage,
              load_p=(0.05, 0.6), sgen_p=(0.03, 0.35), storage_p=(0.05, 0.3))
    return net


def synth_hv(nb=372, n_load=79, n_sgen=37, n_storage=0, n_gen=5, seed=0,
             name="synth_hv") -> pn.Net:
    """Meshed 110-kV grid: bus 0 = 380 kV slack feeding six 110-kV substations
    through 300-MVA transformers; backbone tree plus ~30 % loop-closing lines."""
    rng = np.random.default_rng(seed)
    net = pn.create_empty_network(name=name, sn_mva=1.0)
    pn.create_buses(net, 1, 380.0)
    pn.create_buses(net, nb - 1, 110.0)
    pn.create_ext_grid(net, 0, vm_pu=1.0)
    n110 = nb - 1
    xy = rng.uniform(0.0, 1.0, (n110, 2))
    order = np.argsort(xy[:, 0] + 0.3 * xy[:, 1])
    xy = xy[order]
    fb, tb = [], []
    for i in range(1, n110):
        d = np.hypot(*(xy[:i] - xy[i]).T)
        j = int(np.argmin(d))
        fb.append(1 + j)
        tb.append(1 + i)
    have = set(zip(fb, tb))
    n_extra = int(0.32 * n110)
    tries = 0
    while n_extra > 0 and tries < 100000:
        tries += 1
        i = int(rng.integers(0, n110))
        d = np.hypot(*(xy - xy[i]).T)
        d[i] = np.inf
        cand = np.argsort(d)[:4]
        j = int(cand[rng.integers(0, len(cand))])
        a, b = 1 + min(i, j), 1 + max(i, j)
        if (a, b) in have:
            continue
        have.add((a, b))
        fb.append(a)
        tb.append(b)
        n_extra -= 1
    n_line = len(fb)
    fbv, tbv = np.array(fb), np.array(tb)
    length = 4.0 + 60.0 * np.hypot(*(xy[fbv - 1] - xy[tbv - 1]).T)
    typ = rng.integers(0, 2, n_line)
    par = np.array(_HV_LINES)[typ]
    pn.create_lines_from_parameters(net, fb, tb, length, par[:, 0], par[:, 1],
                                    par[:, 2], par[:, 3], parallel=2.0)
    # EHV/HV injection points spread over the area
    anchors = np.array([[0.2, 0.25], [0.5, 0.2], [0.8, 0.3], [0.25, 0.75], [0.55, 0.8], [0.8, 0.7]])
    centre = sorted({int(np.argmin(np.hypot(*(xy - a).T))) for a in anchors})
    for c in centre:
        pn.create_transformer_from_parameters(
            net, 0, 1 + int(c), sn_mva=300.0, vn_hv_kv=380.0, vn_lv_kv=110.0,
            vkr_percent=0.25, vk_percent=14.0, pfe_kw=100.0, i0_percent=0.06,
            shift_degree=0.0, tap_side="hv", tap_neutral=0.0, tap_min=-9.0,
            tap_max=9.0, tap_step_percent=1.5, tap_pos=0.0)
    hv_buses = np.arange(1, nb)
    _populate(net, rng, hv_buses, n_load, n_sgen, n_storage,
              load_p=(3.0, 12.0), sgen_p=(5.0, 20.0), storage_p=(2.0, 8.0))
    if n_gen:
        gb = rng.choice(hv_buses, n_gen, replace=False)
        for b in gb:
            pn.create_gen(net, int(b), p_mw=float(rng.uniform(20.0, 60.0)), vm_pu=1.0)
    return net


def _populate(net, rng, buses, n_load, n_sgen, n_storage, load_p, sgen_p, storage_p):
    def pick(n):
        if n <= len(buses):
            return rng.choice(buses, n, replace=False)
        return np.concatenate([rng.permutation(buses),
                               rng.choice(buses, n - len(buses), replace=True)])
    if n_load:
        p = rng.uniform(*load_p, n_load)
        cosphi = rng.uniform(0.93, 0.97, n_load)
        pn.create_loads(net, pick(n_load), p, p * np.tan(np.arccos(cosphi)))
    if n_sgen:
        pn.create_sgens(net, pick(n_sgen), rng.uniform(*sgen_p, n_sgen), 0.0)
    if n_storage:
        pn.create_storages(net, pick(n_storage), rng.uniform(*storage_p, n_storage), 0.0)


# -------------------------------------------------------------------------- profiles
def synth_profiles(net, n_steps=N_SIMBENCH_STEPS, seed=0) -> dict:
    """Quarter-hour absolute-value profiles keyed like
    ``simbench.get_absolute_values(net, True)``: ``(table, column) -> DataFrame
    [n_steps, n_units]`` whose columns are the unit indices."""
    rng = np.random.default_rng(seed + 12345)
    t = np.arange(n_steps)
    day = 2 * np.pi * (t % 96) / 96.0
    week = 2 * np.pi * (t % 672) / 672.0
    year = 2 * np.pi * t / float(N_SIMBENCH_STEPS)

    def smooth_noise(n_units, scale, knots):
        k = max(2, n_steps // knots + 2)
        pts = rng.normal(0.0, scale, (k, n_units))
        x = np.linspace(0, k - 1, n_steps)
        i0 = np.floor(x).astype(int).clip(0, k - 2)
        w = (x - i0)[:, None]
        return pts[i0] * (1 - w) + pts[i0 + 1] * w

    # simbench returns every key, with zero columns for absent unit types
    out = {key: pd.DataFrame(index=np.arange(n_steps), columns=net[key[0]].index, dtype=float)
           for key in (("load", "p_mw"), ("load", "q_mvar"), ("sgen", "p_mw"),
                       ("gen", "p_mw"), ("storage", "p_mw"))}
    nl = len(net.load)
    if nl:
        base = (0.55 + 0.22 * np.sin(day - 2.2) + 0.12 * np.sin(2 * day - 0.6)
                + 0.06 * np.cos(week) + 0.08 * np.cos(year))[:, None]
        shape = np.clip(base * (1 + smooth_noise(nl, 0.12, 16)) + rng.normal(0, 0.02, (n_steps, nl)),
                        0.08, 1.0)
        shape /= shape.max(axis=0, keepdims=True)
        p_nom = net.load.p_mw.to_numpy(float)
        q_nom = net.load.q_mvar.to_numpy(float)
        qshape = np.clip(shape * (1 + smooth_noise(nl, 0.05, 48)), 0.05, 1.0)
        out[("load", "p_mw")] = pd.DataFrame(shape * p_nom, columns=net.load.index)
        out[("load", "q_mvar")] = pd.DataFrame(qshape * q_nom, columns=net.load.index)
    ns = len(net.sgen)
    if ns:
        is_pv = rng.random(ns) < 0.7
        sun = np.clip(np.sin(day - np.pi / 2) * 1.15 + 0.25 * np.cos(year + np.pi), 0.0, None)[:, None]
        pv = np.clip(sun * (0.75 + smooth_noise(ns, 0.25, 8)), 0.0, 1.0)
        wind = np.clip(0.35 + smooth_noise(ns, 0.3, 40) + 0.1 * np.cos(year)[:, None], 0.0, 1.0)
        shape = np.where(is_pv[None, :], pv, wind)
        shape /= np.maximum(shape.max(axis=0, keepdims=True), 1e-9)
        out[("sgen", "p_mw")] = pd.DataFrame(
            shape * net.sgen.p_mw.to_numpy(float), columns=net.sgen.index)
    nst = len(net.storage)
    if nst:
        shape = np.clip(0.6 * np.sin(day[:, None] + rng.uniform(0, 2 * np.pi, nst)[None, :])
                        + smooth_noise(nst, 0.3, 12), -1.0, 1.0)
        shape /= np.abs(shape).max(axis=0, keepdims=True)
        out[("storage", "p_mw")] = pd.DataFrame(
            shape * net.storage.p_mw.to_numpy(float), columns=net.storage.index)
    ng = len(net.gen)
    if ng:
        shape = np.clip(0.6 + smooth_noise(ng, 0.2, 96), 0.1, 1.0)
        shape /= shape.max(axis=0, keepdims=True)
        out[("gen", "p_mw")] = pd.DataFrame(
            shape * net.gen.p_mw.to_numpy(float), columns=net.gen.index)
    return out


# ------------------------------------------------ reference net preparation (L1, a12)
def set_unit_scaling(net, gen_scaling=1.0, load_scaling=1.0, storage_scaling=1.0):
    """Reference ``opfgym/simbench/build_simbench_net.py:26-31``."""
    net.sgen["scaling"] = float(gen_scaling)
    net.gen["scaling"] = float(gen_scaling)
    net.load["scaling"] = float(load_scaling)
    net.storage["scaling"] = float(storage_scaling)


def set_system_constraints(net, voltage_band=None, max_loading=None):
    """Reference ``build_simbench_net.py:34-42``."""
    if voltage_band:
        net.bus["max_vm_pu"] = 1.0 + voltage_band
        net.bus["min_vm_pu"] = 1.0 - voltage_band
    if max_loading:
        net.line["max_loading_percent"] = float(max_loading)
        net.trafo["max_loading_percent"] = float(max_loading)


def repair_profiles(net, profiles):
    """Reference ``build_simbench_net.py:45-64``: clamp negative sgen power and
    drop units whose profile is constant."""
    if ("sgen", "p_mw") in profiles:
        profiles[("sgen", "p_mw")] = profiles[("sgen", "p_mw")].clip(lower=0.0)
    for key in list(profiles.keys()):
        df = profiles[key]
        const = (df.max(axis=0) == df.min(axis=0)).to_numpy()
        if const.any():
            table = net[key[0]]
            net[key[0]] = table.drop(table.index[const])
            profiles[key] = df.drop(columns=df.columns[const])


def set_constraints_from_profiles(net, profiles):
    """Reference ``build_simbench_net.py:67-97``."""
    for (unit_type, column), prof in profiles.items():
        df = net[unit_type]
        scaling = df.scaling.to_numpy(float)
        pmax = prof.max(axis=0).to_numpy()
        pmin = prof.min(axis=0).to_numpy()
        if unit_type == "storage":
            top = np.maximum(np.abs(pmax), np.abs(pmin))
            df[f"max_max_{column}"] = top * scaling
            df[f"min_min_{column}"] = -top * scaling
        else:
            df[f"max_max_{column}"] = pmax * scaling
            df[f"min_min_{column}"] = pmin * scaling
        df[f"mean_{column}"] = prof.mean(axis=0).to_numpy()
        df[f"std_dev_{column}"] = prof.std(axis=0).to_numpy()
    load_p = profiles[("load", "p_mw")].sum(axis=1)
    gen_p = profiles[("sgen", "p_mw")].sum(axis=1) if ("sgen", "p_mw") in profiles else 0.0
    diff = load_p - gen_p
    net.ext_grid["max_max_p_mw"] = float(diff.max())
    net.ext_grid["min_min_p_mw"] = float(diff.min())
    net.ext_grid["mean_p_mw"] = float(diff.mean())
    load_q = profiles[("load", "q_mvar")].sum(axis=1)
    net.ext_grid["max_max_q_mvar"] = float(load_q.max())
    net.ext_grid["min_min_q_mvar"] = float(load_q.min())
    net.ext_grid["mean_q_mvar"] = float(load_q.mean())


def _resize_units(net, table, count_thr_scale):
    """Give the first ``count`` units a nominal power safely above
    ``threshold/scaling`` and push the rest safely below, so that the number of
    controllable units matches the reference benchmark's action count."""
    count, thr, scale = count_thr_scale
    df = net[table]
    if not len(df):
        return
    p = df.p_mw.to_numpy(float).copy()
    lim = thr / scale
    if thr > 0:
        small = np.minimum(p[count:], 0.8 * lim)
        p[count:] = small
        p[:count] = np.maximum(p[:count], 1.25 * lim) + lim * np.linspace(0.1, 1.5, count)
    net[table]["p_mw"] = p


def raw_standin(simbench_network_name, seed=0) -> pn.Net:
    """The synthetic stand-in for ``sb.get_simbench_net(name)``: the bare grid,
    before the reference's scaling/constraint post-processing."""
    spec = dict(_STANDINS[simbench_network_name])
    kind = spec.pop("kind")
    big_sgen = spec.pop("big_sgen")
    big_storage = spec.pop("big_storage")
    big_load = spec.pop("big_load", None)
    net = (synth_mv if kind == "mv" else synth_hv)(seed=seed, name=simbench_network_name, **spec)
    _resize_units(net, "sgen", big_sgen)
    _resize_units(net, "storage", big_storage)
    if big_load:
        _resize_units(net, "load", big_load)
        p = net.load.p_mw.to_numpy(float)
        net.load["q_mvar"] = p * np.tan(np.arccos(0.95))
    # SimBench storage tables carry reactive limits (EcoDispatch lists
    # ('storage','q_mvar') as an observation even when the table is empty)
    net.storage["min_q_mvar"] = 0.0
    net.storage["max_q_mvar"] = 0.0
    return net


def build_simbench_net(simbench_network_name, gen_scaling=1.0, load_scaling=1.0,
                       storage_scaling=1.0, voltage_band=0.05, max_loading=80,
                       n_profile_steps=N_SIMBENCH_STEPS, seed=0, **kwargs):
    """Same signature and post-processing as the reference builder
    (``opfgym/simbench/build_simbench_net.py:5-23``), on the synthetic stand-in
    of the same name (SimBench itself is not installable here).
    Returns ``(net, profiles)``."""
    net = raw_standin(simbench_network_name, seed)
    set_unit_scaling(net, gen_scaling, load_scaling, storage_scaling)
    set_system_constraints(net, voltage_band, max_loading)
    profiles = synth_profiles(net, n_steps=n_profile_steps, seed=seed)
    repair_profiles(net, profiles)
    set_constraints_from_profiles(net, profiles)
    return net, profiles
