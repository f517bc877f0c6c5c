"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  Not imported by the product package.

Independent restatement of pandapower's net -> ppc conversion (``pd2ppc.py``,
``build_bus.py``, ``build_branch.py``, ``build_gen.py`` of pandapower 2.x, the
dependency behind reference ``opfgym/opf_env.py:703``; pinned ``>=2.13.1,<3.0`` in
``/root/reference/pyproject.toml:32``, absent from the image -> restated from its
published algorithm).

It exists so that the oracle does NOT share the product's ``opfgym_b200/ppc.py``
(round-1 verdict: "both arms of every power-flow comparison consume the same
matrices").  Different structure on purpose: element-by-element Python loops over
dictionaries, buses numbered by a breadth-first walk from the slack, complex
shunt admittances carried as one number the way pandapower's complex ``BR_B`` column
does.  ``tests/test_ppc_ref.py`` checks the two builders against each other on every
stand-in grid (after mapping the bus numbering), so a slip in either shows up.

PARITY STATUS: unpinned at the pandapower boundary, like ``oracle/pf.py``.
"""
from __future__ import annotations

import math
from collections import deque
from types import SimpleNamespace

import numpy as np

# PYPOWER column order (idx_bus / idx_gen / idx_brch), + BR_G as in pandapower >= 2.14
BUS_I, BUS_TYPE, PD, QD, GS, BS, BUS_AREA, VM, VA, BASE_KV, ZONE, VMAX, VMIN = range(13)
GEN_BUS, PG, QG, QMAX, QMIN, VG, MBASE, GEN_STATUS, PMAX, PMIN = range(10)
(F_BUS, T_BUS, BR_R, BR_X, BR_B, RATE_A, RATE_B, RATE_C, TAP, SHIFT,
 BR_STATUS, ANGMIN, ANGMAX, BR_G) = range(14)
PQ, PV, REF, NONE = 1, 2, 3, 4


def _rows(df):
    return [(idx, row) for idx, row in zip(df.index, df.to_dict("records"))]


def _num(v, default=0.0):
    try:
        v = float(v)
    except (TypeError, ValueError):
        return default
    return default if math.isnan(v) else v


def build(net, calculate_voltage_angles=True, trafo_model="t"):
    """Returns a namespace with ``base_mva, bus, gen, branch`` (ppc matrices) and the lookups
    ``bus_lookup`` (pandapower bus index -> ppc bus or -1), ``line_branch`` / ``trafo_branch``
    (element index -> branch row or -1), ``ext_grid_gen`` / ``gen_gen``, ``rate_f`` / ``rate_t``."""
    sn = float(net.sn_mva)
    f_hz = float(net.f_hz)

    # ---- buses: fuse closed bus-bus switches (build_bus.py: _build_bus_ppc / bus lookup) ----
    in_service = {int(i): bool(r["in_service"]) for i, r in _rows(net.bus)}
    vn_kv = {int(i): float(r["vn_kv"]) for i, r in _rows(net.bus)}
    rep = {b: b for b in in_service}

    def find(b):
        while rep[b] != b:
            rep[b] = rep[rep[b]]
            b = rep[b]
        return b

    switches = _rows(net.switch) if len(net.switch) else []
    for _, s in switches:
        if s["et"] == "b" and bool(s["closed"]):
            a, b = int(s["bus"]), int(s["element"])
            if in_service[a] and in_service[b]:
                ra, rb = find(a), find(b)
                if ra != rb:
                    rep[max(ra, rb)] = min(ra, rb)

    # ---- branches as (kind, index, node_from, node_to); open line switches (build_branch.py:
    # _switch_branches): one open end -> auxiliary bus, both ends open -> out of service ----
    open_ends = {}
    for _, s in switches:
        if s["et"] == "l" and not bool(s["closed"]):
            open_ends.setdefault(int(s["element"]), set()).add(int(s["bus"]))
    open_trafos = {int(s["element"]) for _, s in switches if s["et"] == "t" and not bool(s["closed"])}
    aux_vn = {}
    edges = []
    for idx, ln in _rows(net.line):
        fb, tb = int(ln["from_bus"]), int(ln["to_bus"])
        if not (bool(ln["in_service"]) and in_service[fb] and in_service[tb]):
            continue
        opened = open_ends.get(int(idx), set())
        if fb in opened and tb in opened:
            continue
        nf, nt = ("bus", find(fb)), ("bus", find(tb))
        if fb in opened:
            nf = ("aux", int(idx)); aux_vn[nf] = vn_kv[fb]
        elif tb in opened:
            nt = ("aux", int(idx)); aux_vn[nt] = vn_kv[tb]
        edges.append(("line", int(idx), nf, nt, ln))
    for idx, tr in _rows(net.trafo):
        hb, lb = int(tr["hv_bus"]), int(tr["lv_bus"])
        if not (bool(tr["in_service"]) and in_service[hb] and in_service[lb]) or int(idx) in open_trafos:
            continue
        edges.append(("trafo", int(idx), ("bus", find(hb)), ("bus", find(lb)), tr))
    for idx, im in (_rows(net.impedance) if len(getattr(net, "impedance", ())) else []):
        fb, tb = int(im["from_bus"]), int(im["to_bus"])
        if bool(im["in_service"]) and in_service[fb] and in_service[tb]:
            edges.append(("impedance", int(idx), ("bus", find(fb)), ("bus", find(tb)), im))

    # ---- connectivity (pd2ppc.py: _check_connectivity): only what hangs on a slack survives;
    # the walk order IS this builder's bus numbering ----
    adjacency = {}
    for _, _, a, b, _ in edges:
        adjacency.setdefault(a, []).append(b)
        adjacency.setdefault(b, []).append(a)
    slack_nodes = []
    for _, eg in _rows(net.ext_grid):
        if bool(eg["in_service"]) and in_service[int(eg["bus"])]:
            slack_nodes.append(("bus", find(int(eg["bus"]))))
    if len(net.gen) and "slack" in net.gen.columns:
        for _, g in _rows(net.gen):
            if bool(g["slack"]) and bool(g["in_service"]) and in_service[int(g["bus"])]:
                slack_nodes.append(("bus", find(int(g["bus"]))))
    number = {}
    queue = deque()
    for s in slack_nodes:
        if s not in number:
            number[s] = len(number)
            queue.append(s)
    while queue:
        a = queue.popleft()
        for b in adjacency.get(a, ()):
            if b not in number:
                number[b] = len(number)
                queue.append(b)
    nb = len(number)
    base_kv = np.zeros(nb)
    for node, k in number.items():
        base_kv[k] = vn_kv[node[1]] if node[0] == "bus" else aux_vn[node]

    def ppc_bus(pp_bus):
        pp_bus = int(pp_bus)
        if not in_service[pp_bus]:
            return -1
        return number.get(("bus", find(pp_bus)), -1)

    bus = np.zeros((nb, 13))
    bus[:, BUS_I] = np.arange(nb)
    bus[:, BUS_TYPE] = PQ
    bus[:, BUS_AREA] = bus[:, ZONE] = 1
    bus[:, BASE_KV] = base_kv
    bus[:, VMAX], bus[:, VMIN] = 2.0, 0.0

    # ---- bus injections (build_bus.py: _calc_pq_elements_and_add_on_ppc): p*scaling*in_service ----
    for table, sign in (("load", +1.0), ("sgen", -1.0), ("storage", +1.0)):
        for _, el in _rows(net[table]):
            k = ppc_bus(el["bus"])
            if k < 0 or not bool(el["in_service"]):
                continue
            w = sign * float(el["scaling"])
            bus[k, PD] += w * float(el["p_mw"])
            bus[k, QD] += w * _num(el["q_mvar"])
    for _, sh in (_rows(net.shunt) if len(net.shunt) else []):   # _calc_shunts_and_add_on_ppc
        k = ppc_bus(sh["bus"])
        if k < 0 or not bool(sh["in_service"]):
            continue
        v_ratio = (base_kv[k] / float(sh["vn_kv"])) ** 2
        bus[k, GS] += float(sh["p_mw"]) * float(sh["step"]) * v_ratio
        bus[k, BS] -= float(sh["q_mvar"]) * float(sh["step"]) * v_ratio

    for _, w in (_rows(net.ward) if len(getattr(net, "ward", ())) else []):
        # a ward = constant power (joins PD/QD) + constant impedance rated at the bus voltage (joins GS/BS)
        k = ppc_bus(w["bus"])
        if k < 0 or not bool(w["in_service"]):
            continue
        bus[k, PD] += float(w["ps_mw"])
        bus[k, QD] += float(w["qs_mvar"])
        bus[k, GS] += float(w["pz_mw"])
        bus[k, BS] -= float(w["qz_mvar"])

    # ---- generators (build_gen.py): ext_grid rows (REF), then gen rows (PV) ----
    gen_rows, ext_grid_gen, gen_gen = [], {}, {}
    for idx, eg in _rows(net.ext_grid):
        k = ppc_bus(eg["bus"])
        ext_grid_gen[int(idx)] = -1
        if k < 0 or not bool(eg["in_service"]):
            continue
        row = np.zeros(10)
        row[GEN_BUS], row[VG], row[MBASE], row[GEN_STATUS] = k, float(eg["vm_pu"]), sn, 1
        ext_grid_gen[int(idx)] = len(gen_rows)
        gen_rows.append(row)
        bus[k, BUS_TYPE] = REF
        bus[k, VA] = _num(eg.get("va_degree", 0.0))
    for idx, g in (_rows(net.gen) if len(net.gen) else []):
        k = ppc_bus(g["bus"])
        gen_gen[int(idx)] = -1
        if k < 0 or not bool(g["in_service"]):
            continue
        row = np.zeros(10)
        row[GEN_BUS], row[MBASE], row[GEN_STATUS] = k, sn, 1
        row[PG] = float(g["p_mw"]) * float(g["scaling"])
        row[VG] = float(g["vm_pu"])
        row[QMAX] = _num(g.get("max_q_mvar"), 1e9)
        row[QMIN] = _num(g.get("min_q_mvar"), -1e9)
        gen_gen[int(idx)] = len(gen_rows)
        gen_rows.append(row)
        if bool(g.get("slack", False)):
            bus[k, BUS_TYPE] = REF
        elif bus[k, BUS_TYPE] == PQ:
            bus[k, BUS_TYPE] = PV
    gen = np.array(gen_rows).reshape(-1, 10)

    # initial |V| (pd2ppc / _init_runpp_options, init='auto'): mean set-point; controlled buses at their own
    setpoints = [float(eg["vm_pu"]) for _, eg in _rows(net.ext_grid)]
    setpoints += [float(g["vm_pu"]) for _, g in (_rows(net.gen) if len(net.gen) else [])]
    init_vm = sum(setpoints) / len(setpoints) if setpoints else 1.0
    bus[:, VM] = init_vm
    for row in gen_rows:
        bus[int(row[GEN_BUS]), VM] = row[VG]

    # ---- branch table (build_branch.py) ----
    branch_rows, rate_f, rate_t = [], [], []
    line_branch = {int(i): -1 for i in net.line.index}
    trafo_branch = {int(i): -1 for i in net.trafo.index}
    impedance_branch = {int(i): -1 for i in (net.impedance.index if len(getattr(net, "impedance", ())) else [])}
    # pandapower stacks the branch table by element type: lines, transformers, impedances
    edges.sort(key=lambda e: ("line", "trafo", "impedance").index(e[0]))
    for kind, idx, a, b, el in edges:
        if a not in number or b not in number:
            continue
        fa, tb = number[a], number[b]
        row = np.zeros(14)
        row[F_BUS], row[T_BUS], row[TAP], row[BR_STATUS] = fa, tb, 1.0, 1.0
        row[ANGMIN], row[ANGMAX] = -360.0, 360.0
        if kind == "line":                                   # _calc_line_parameter
            length, parallel = float(el["length_km"]), float(el["parallel"])
            base_r = base_kv[fa] ** 2 / sn
            row[BR_R] = float(el["r_ohm_per_km"]) * length / base_r / parallel
            row[BR_X] = float(el["x_ohm_per_km"]) * length / base_r / parallel
            row[BR_B] = 2.0 * math.pi * f_hz * float(el["c_nf_per_km"]) * 1e-9 * base_r * length * parallel
            row[BR_G] = _num(el.get("g_us_per_km")) * 1e-6 * base_r * length * parallel
            i_max = float(el["max_i_ka"]) * float(el["df"]) * parallel
            row[RATE_A] = i_max * base_kv[fa] * math.sqrt(3.0)
            # results_branch.py: i_ka = |S| / (sqrt3 * vm * vn_kv); loading = i_ka / (max_i_ka * df * parallel)
            rate_f.append(1.0 / (math.sqrt(3.0) * base_kv[fa] * i_max))
            rate_t.append(1.0 / (math.sqrt(3.0) * base_kv[tb] * i_max))
            line_branch[idx] = len(branch_rows)
        elif kind == "impedance":                            # _calc_impedance_parameter
            scale = sn / float(el["sn_mva"])
            z_ft = complex(float(el["rft_pu"]), float(el["xft_pu"])) * scale
            z_tf = complex(float(el["rtf_pu"]), float(el["xtf_pu"])) * scale
            if z_ft != z_tf:
                raise NotImplementedError("asymmetric impedance element")
            row[BR_R], row[BR_X] = z_ft.real, z_ft.imag
            rate_f.append(0.0)
            rate_t.append(0.0)
            impedance_branch[idx] = len(branch_rows)
        else:                                                # _calc_branch_values_from_trafo_df
            vn_hv_bus, vn_lv_bus = base_kv[fa], base_kv[tb]
            vn_hv, vn_lv = float(el["vn_hv_kv"]), float(el["vn_lv_kv"])
            parallel, sn_t = float(el["parallel"]), float(el["sn_mva"])
            # _calc_tap_from_dataframe (ratio tap changer; tap_step_degree = 0)
            step = _num(el.get("tap_step_percent")) * (_num(el.get("tap_pos")) - _num(el.get("tap_neutral"))) / 100.0
            vn_hv_tap, vn_lv_tap = vn_hv, vn_lv
            if el.get("tap_side") == "hv":
                vn_hv_tap = vn_hv * (1.0 + step)
            elif el.get("tap_side") == "lv":
                vn_lv_tap = vn_lv * (1.0 + step)
            # _calc_nominal_ratio_from_dataframe
            row[TAP] = (vn_hv_tap / vn_lv_tap) / (vn_hv_bus / vn_lv_bus)
            row[SHIFT] = _num(el.get("shift_degree")) if calculate_voltage_angles else 0.0
            # _calc_r_x_from_dataframe: short-circuit impedance on the LV side
            tap_lv = (vn_lv_tap / vn_lv_bus) ** 2 * sn
            z_sc = float(el["vk_percent"]) / 100.0 / sn_t * tap_lv
            r_sc = float(el["vkr_percent"]) / 100.0 / sn_t * tap_lv
            x_sc = math.copysign(math.sqrt(z_sc ** 2 - r_sc ** 2), z_sc)
            z_series = complex(r_sc, x_sc) / parallel
            # _calc_y_from_dataframe: magnetising branch, kept as ONE complex admittance g - jb
            base_r = vn_lv_bus ** 2 / sn
            pfe = float(el["pfe_kw"]) * 1e-3
            i0 = float(el["i0_percent"])
            g_m = pfe / vn_lv ** 2 * base_r
            b_sq = max((i0 / 100.0 * sn_t) ** 2 - pfe ** 2, 0.0)
            b_m = math.sqrt(b_sq) * base_r / vn_lv ** 2 * (0.0 if i0 == 0 else math.copysign(1.0, i0))
            y_m = complex(g_m, -b_m) / (vn_lv_tap / vn_lv) ** 2 * parallel
            if trafo_model == "t" and y_m != 0:               # _wye_delta: T -> pi
                za, zc = z_series / 2.0, 1.0 / y_m
                z_sum = za * za + 2.0 * za * zc
                z_series = z_sum / zc
                y_m = 2.0 / (z_sum / za)
            row[BR_R], row[BR_X] = z_series.real, z_series.imag
            row[BR_G], row[BR_B] = y_m.real, y_m.imag
            df = float(el["df"])
            row[RATE_A] = sn_t * parallel * df
            # trafo_loading='current': 100 * max(i_hv * vn_hv, i_lv * vn_lv) * sqrt3 / sn / (df * parallel)
            rate_f.append(vn_hv / (vn_hv_bus * sn_t * parallel * df))
            rate_t.append(vn_lv / (vn_lv_bus * sn_t * parallel * df))
            trafo_branch[idx] = len(branch_rows)
        branch_rows.append(row)
    branch = np.array(branch_rows).reshape(-1, 14)

    as_array = lambda d, index: np.array([d[int(i)] for i in index], dtype=np.int64)
    return SimpleNamespace(
        base_mva=sn, bus=bus, gen=gen, branch=branch, init_vm_pu=init_vm,
        bus_lookup=np.array([ppc_bus(i) for i in net.bus.index], dtype=np.int64),
        line_branch=as_array(line_branch, net.line.index), trafo_branch=as_array(trafo_branch, net.trafo.index),
        ext_grid_gen=as_array(ext_grid_gen, net.ext_grid.index),
        gen_gen=as_array(gen_gen, net.gen.index) if len(net.gen) else np.zeros(0, np.int64),
        impedance_branch=as_array(impedance_branch, net.impedance.index) if impedance_branch
        else np.zeros(0, np.int64),
        rate_f=np.array(rate_f), rate_t=np.array(rate_t))


class RefBuilder:
    """Object form with the ``build(net)`` method ``oracle.pf.runpp`` expects from a builder."""

    def __init__(self, net=None, **kwargs):
        self.kwargs = kwargs

    def build(self, net):
        return build(net, **self.kwargs)
