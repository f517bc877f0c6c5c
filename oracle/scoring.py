"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  Not imported by the product package.

Single-environment NumPy restatement of the reference's scoring layer, evaluated
on a ``Net`` whose ``res_*`` tables have been filled (by ``oracle.pf.runpp`` or by
hand, as the reference's own tests do):

* ``pandapower_costs``   <- reference ``opfgym/objective.py:6-87``
* ``violation_metrics``  <- reference ``opfgym/constraints.py:70-128``
* ``reward`` / ``cost``  <- reference ``opfgym/reward.py:61-98`` and subclasses
* ``step_reward``        <- reference ``opfgym/opf_env.py:493-530``

PARITY STATUS: pinned.  ``tests/test_oracle_scoring.py`` replays every
known-answer case of the reference's ``tests/test_constraints.py``,
``tests/test_objective.py`` and ``tests/test_reward.py`` (SURVEY.md App.
C.1-C.3), and ``tests/golden/scoring_*.npz`` holds outputs of the reference's
*own* modules run in the build container on random tables
(``tests/golden/make_golden.py``).
"""
from __future__ import annotations

import numpy as np


# ------------------------------------------------------------------- objective
def poly_costs(net) -> np.ndarray:
    """``get_polynomial_costs`` (:34-45): [p-costs of all rows, q-costs of all rows]."""
    pc = net.poly_cost
    p = np.array([net["res_" + et].p_mw.loc[el] for et, el in zip(pc.et, pc.element)], float)
    q = np.array([net["res_" + et].q_mvar.loc[el] for et, el in zip(pc.et, pc.element)], float)
    cost_p = (pc.cp0_eur.to_numpy(float) + pc.cp1_eur_per_mw.to_numpy(float) * p
              + pc.cp2_eur_per_mw2.to_numpy(float) * p ** 2)
    cost_q = (pc.cq0_eur.to_numpy(float) + pc.cq1_eur_per_mvar.to_numpy(float) * q
              + pc.cq2_eur_per_mvar2.to_numpy(float) * q ** 2)
    return np.r_[cost_p, cost_q]


def pwl_costs(net) -> np.ndarray:
    """``get_piecewise_linear_costs`` (:57-77).  Segments are walked index-wise
    across all rows (so the walk stops at the shortest row); ``outside`` has no
    sign test."""
    pw = net.pwl_cost
    power = np.array([
        net["res_" + et]["p_mw" if pt == "p" else "q_mvar"].loc[el]
        for et, el, pt in zip(pw.et, pw.element, pw.power_type)], float)
    n_seg = min(len(pts) for pts in pw.points)
    total = np.zeros(len(pw))
    sgn = np.sign(power)
    mag = np.abs(power)
    for s in range(n_seg):
        seg = np.array([pts[s] for pts in pw.points], float)
        lo, hi, price = seg[:, 0], seg[:, 1], seg[:, 2]
        near = np.minimum(np.abs(lo), np.abs(hi))
        far = np.maximum(np.abs(lo), np.abs(hi))
        beyond = mag > far
        within = (mag > near) & (sgn == np.sign(lo + hi)) & ~beyond
        total += np.where(beyond, sgn * (hi - lo) * price, 0.0)
        total += np.where(within, sgn * (mag - near) * price, 0.0)
    return total


def pandapower_costs(net) -> np.ndarray:
    parts = []
    if len(net.poly_cost):
        parts.append(poly_costs(net))
    if len(net.pwl_cost):
        parts.append(pwl_costs(net))
    return np.concatenate(parts) if parts else np.array([])


# ----------------------------------------------------------------- constraints
def violation_metrics(constraint, net) -> dict:
    """``Constraint.get_violation_metrics`` (:70-88) for a product-side
    constraint record (``opfgym_b200.constraints``)."""
    c = constraint
    values = net["res_" + c.unit_type][c.values_column].to_numpy(float) * c.value_scale
    table = net[c.unit_type]
    mult = c.boundary_multiplier(net)
    violation, count = 0.0, 0
    for side in ("min", "max"):
        col = f"{side}_{c.values_column}"
        if col not in table.columns:
            continue
        bound = table[col].to_numpy(float) * mult
        with np.errstate(invalid="ignore"):
            bad = values > bound if side == "max" else values < bound
        n_bad = int(bad.sum())
        count += n_bad
        if n_bad:
            excess = np.abs(values - bound)[bad]
            violation += excess.max() if c.only_worst_case_violations else excess.sum()
    factor = c.autoscale_factor(net)
    if c.autoscale_violation:
        violation *= factor
    penalty = -(violation ** c.penalty_power * c.penalty_factor
                + count * c.violation_count_penalty)
    return {"valid": count == 0, "violation": violation, "penalty": penalty}


# ---------------------------------------------------------------------- reward
def reward(rf, objective: float, penalty: float, valid: bool) -> float:
    """``RewardFunction.__call__`` (:61-73) from the flat parameter record."""
    p = rf.device_params()
    kind = p["kind"]
    obj, pen = objective, penalty
    if kind == 1:      # Replacement :246-251
        obj = obj + p["valid_reward"] if valid else 0.0
    elif kind == 2:    # Parameterized :289-299
        pen = pen + p["valid_reward"] if valid else pen - p["invalid_penalty"]
        obj = obj if valid else obj * p["invalid_objective_share"]
    elif kind == 3:    # OnlyObjective :313-320
        pen = 0.0
    obj = obj * p["objective_factor"] + p["objective_bias"]
    pen = pen * p["penalty_factor"] + p["penalty_bias"]
    w = p["penalty_weight"]
    r = obj + pen if np.isnan(w) else obj * (1 - w) + pen * w
    if not np.isnan(p["clip_lo"]):
        r = min(max(r, p["clip_lo"]), p["clip_hi"])
    return r


def cost(rf, penalty: float, valid: bool) -> float:
    """``calculate_cost`` (:93-98, Parameterized override :301-305)."""
    p = rf.device_params()
    if valid:
        return 0.0
    c = abs(penalty * p["penalty_factor"])
    return c + p["invalid_penalty"] if p["kind"] == 2 else c


def step_reward(net, constraints, rf, initial_obj=None) -> dict:
    """``OpfEnv.calculate_reward`` (opf_env.py:515-530)."""
    obj_arr = -pandapower_costs(net)
    if initial_obj is not None:
        obj_arr = obj_arr - initial_obj
    objective = float(np.sum(obj_arr))
    metrics = [violation_metrics(c, net) for c in constraints]
    valids = np.array([m["valid"] for m in metrics], bool)
    violations = np.array([m["violation"] for m in metrics], float)
    penalties = np.array([m["penalty"] for m in metrics], float)
    penalty = float(np.sum(penalties))
    valid = bool(valids.all())
    return {"reward": reward(rf, objective, penalty, valid),
            "cost": cost(rf, penalty, valid), "objective": objective,
            "penalty": penalty, "valid": valid, "valids": valids,
            "violations": violations, "unscaled_penalties": penalties}
