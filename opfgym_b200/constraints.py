"""Constraint *descriptions* for the batched engine.

Same class names and keyword arguments as reference ``opfgym/constraints.py``
(``Constraint`` :19-128, the five default subclasses :131-192 and
``create_default_constraints`` :195-238), but a constraint here is a plain
record: it names the result column it bounds and carries the violation/penalty
parameters.  ``opfgym_b200.compiler`` turns a list of these records into flat
device tables, and the fused scoring kernel (``csrc/score.cuh``) evaluates all
of them for every environment in one launch.  There is no per-constraint
Python in ``step``.

Semantics reproduced on the device:
  invalid = value > max  |  value < min           (NaN compares False)
  violation = sum|value - bound| over invalid (or the max if worst-case)
  violation *= autoscale_violation (if truthy)
  penalty = -(violation**penalty_power * penalty_factor + n_invalid * violation_count_penalty)
  valid = (n_invalid == 0)
"""
from __future__ import annotations

import numpy as np
import pandas as pd


class Constraint:
    def __init__(self, unit_type: str, values_column: str,
                 get_values=None, get_boundaries=None,
                 only_worst_case_violations: bool = False,
                 autoscale_violation=True,
                 scale_bounded_values: bool = False,
                 penalty_factor: float = 1.0,
                 penalty_power: float = 1.0,
                 violation_count_penalty: float = 0.0,
                 value_scale: float = 1.0):
        # Reference constraints.py:27-64: ``get_values(net)`` / ``get_boundaries(net)`` are Python callables on one
        # pandapower net.  Here they are BATCHED callables on the env (the contract of ``objective_function=``):
        #   get_values(env)     -> tensor [num_envs, n] on env.device (read cells with env.col(table, column))
        #   get_boundaries(env) -> {'min' | 'max': tensor [num_envs, n] | [n] | scalar}   (not scaled again, as in
        #                          the reference, where a custom get_boundaries replaces scale_boundary too)
        # Such a constraint does not enter the fused scoring kernel: the env evaluates it with tensor ops behind
        # kernel 5 and recombines penalty, validity, reward and cost (opf_env.BatchedOpfEnv._apply_plugins).
        if (get_values is None) != (get_boundaries is None):
            raise ValueError("a callable constraint needs both get_values and get_boundaries "
                             "(the kernel's own constraints read unit_type / values_column instead)")
        self.get_values_fn = get_values
        self.get_boundaries_fn = get_boundaries
        self.unit_type = unit_type
        self.values_column = values_column
        self.only_worst_case_violations = bool(only_worst_case_violations)
        self.autoscale_violation = autoscale_violation
        self.scale_bounded_values = bool(scale_bounded_values)
        self.penalty_factor = float(penalty_factor)
        self.penalty_power = float(penalty_power)
        self.violation_count_penalty = float(violation_count_penalty)
        # value_scale lets the reference's custom-constraint example
        # (values = res_sgen.p_mw / 2, tests/test_constraints.py:131-147) be
        # expressed without a callable.
        self.value_scale = float(value_scale)

    @property
    def is_batched_callable(self) -> bool:
        return self.get_values_fn is not None

    def batched_metrics(self, env):
        """``get_violation_metrics`` (reference constraints.py:70-88) of a callable constraint for every environment:
        returns (valid[B] bool, violation[B], penalty[B]) as tensors on ``env.device``."""
        xp = env.xp
        values = xp.as_tensor(self.get_values_fn(env), device=env.device).to(xp.float64)
        if values.dim() == 1:
            values = values.reshape(env.num_envs, -1)
        violation = xp.zeros(env.num_envs, dtype=xp.float64, device=env.device)
        n_violations = xp.zeros(env.num_envs, dtype=xp.float64, device=env.device)
        for min_or_max, boundary in self.get_boundaries_fn(env).items():
            if min_or_max not in ("min", "max"):
                raise KeyError(f"get_boundaries returned the key {min_or_max!r}: expected 'min' / 'max'")
            bound = xp.as_tensor(boundary, device=env.device).to(xp.float64)
            invalid = values > bound if min_or_max == "max" else values < bound          # :110-111 (NaN compares false)
            n_violations = n_violations + invalid.sum(dim=1)
            absolute = xp.where(invalid, (values - bound).abs(), xp.zeros_like(values))   # :113-122
            violation = violation + (absolute.max(dim=1).values if self.only_worst_case_violations
                                     else absolute.sum(dim=1))
        if self.autoscale_violation:                                                      # :82-83 (True multiplies by one)
            violation = violation * float(self.autoscale_violation)
        penalty = -(violation ** self.penalty_power * self.penalty_factor
                    + n_violations * self.violation_count_penalty)                        # :124-128
        return n_violations == 0, violation, penalty

    # the factor the kernel multiplies the summed violation with
    def autoscale_factor(self, net) -> float:
        a = self.autoscale_violation
        if a is True:
            return 1.0
        if not a:
            return 1.0  # falsy -> reference skips the multiplication (:82-83)
        return float(a)

    def boundary_multiplier(self, net) -> np.ndarray | float:
        """Reference ``scale_boundary`` :104-108."""
        table = net[self.unit_type]
        if self.scale_bounded_values or ("scaling" in table.columns
                                         and self.values_column in ("p_mw", "q_mvar")):
            return table.scaling.to_numpy(float)
        return 1.0

    def __repr__(self):
        return (f"{type(self).__name__}({self.unit_type}.{self.values_column}, "
                f"autoscale={self.autoscale_violation})")


class VoltageConstraint(Constraint):
    def __init__(self, autoscale_violation=True, **kw):
        if autoscale_violation is True:
            autoscale_violation = 20  # reference :133-135
        super().__init__("bus", "vm_pu", autoscale_violation=autoscale_violation, **kw)


class _OverloadConstraint(Constraint):
    _unit = None

    def __init__(self, autoscale_violation=True, **kw):
        if autoscale_violation is True:
            autoscale_violation = 1 / 30  # reference :144-146,155-157,166-168
        super().__init__(self._unit, "loading_percent",
                         autoscale_violation=autoscale_violation, **kw)


class LineOverloadConstraint(_OverloadConstraint):
    _unit = "line"


class TrafoOverloadConstraint(_OverloadConstraint):
    _unit = "trafo"


class Trafo3wOverloadConstraint(_OverloadConstraint):
    _unit = "trafo3w"


class _ExtGridConstraint(Constraint):
    _column = None

    def __init__(self, **kw):
        super().__init__("ext_grid", self._column, **kw)

    def autoscale_factor(self, net) -> float:
        # Reference :179-182/:189-192: only a *falsy* user value triggers the
        # 1/|sum(mean)| normalisation (A.6 quirk 3); the default True means x1.
        if not self.autoscale_violation:
            mean = net.ext_grid["mean_" + self._column].sum()
            self.autoscale_violation = 1.0 / abs(float(mean))
        return super().autoscale_factor(net)


class ExtGridActivePowerConstraint(_ExtGridConstraint):
    _column = "p_mw"


class ExtGridReactivePowerConstraint(_ExtGridConstraint):
    _column = "q_mvar"


def _defined(net, unit_type: str, column: str) -> bool:
    table = net[unit_type]
    if column not in table.columns or len(table) == 0:
        return False
    numeric = pd.to_numeric(table[column], errors="coerce").to_numpy(float)
    return bool(np.isfinite(numeric).any())


def create_default_constraints(net, constraint_kwargs: dict) -> list:
    """Same discovery order and rules as reference :195-238."""
    kw = dict(constraint_kwargs or {})
    found = []
    if _defined(net, "bus", "max_vm_pu") or _defined(net, "bus", "min_vm_pu"):
        found.append(VoltageConstraint(**kw))
    for cls in (LineOverloadConstraint, TrafoOverloadConstraint, Trafo3wOverloadConstraint):
        if cls._unit in net and _defined(net, cls._unit, "max_loading_percent"):
            found.append(cls(**kw))
    for cls in (ExtGridActivePowerConstraint, ExtGridReactivePowerConstraint):
        if _defined(net, "ext_grid", "max_" + cls._column) or \
                _defined(net, "ext_grid", "min_" + cls._column):
            found.append(cls(**kw))
    return found
