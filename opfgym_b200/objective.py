"""Objective functions of the batched env -- the counterpart of reference ``opfgym/objective.py``.

In the default path nothing here runs per step: the net's ``poly_cost`` / ``pwl_cost`` tables are
flattened by the compiler into device tables and kernel 5 evaluates the reference's cost vector
``[p-costs of all poly rows, q-costs of all poly rows, pwl costs]`` (``objective.py:6-87``, SURVEY.md
App. A.3) for every environment.

This module is the plug-in side, ``BatchedOpfEnv(objective_function=...)`` (reference contract
``opf_env.py:52, 80-84``, batched: the callable receives the env, not a pandapower net):

* ``get_pandapower_costs(env)`` -- the same cost vector as tensor ops, ``[num_envs, 2 n_poly + n_pwl]``,
  in the reference's order.  Passing it as ``objective_function`` reproduces the built-in objective;
  custom objectives usually start from it (``lambda env: get_pandapower_costs(env).sum(1) + extra(env)``).
* ``get_polynomial_costs`` / ``get_piecewise_linear_costs`` -- its two halves (``objective.py:34-45, 57-77``).
* ``cost_vector_layout(net)`` -- which element each entry of the vector belongs to.
"""
from __future__ import annotations


def cost_vector_layout(net) -> list[tuple[str, int, str]]:
    """``[(et, element, 'p'|'q'|'pwl'), ...]`` in the order of the reference's
    ``get_pandapower_costs`` output."""
    pc, pw = net.poly_cost, net.pwl_cost
    out = [(et, int(el), "p") for et, el in zip(pc.et, pc.element)]
    out += [(et, int(el), "q") for et, el in zip(pc.et, pc.element)]
    out += [(et, int(el), "pwl") for et, el in zip(pw.et, pw.element)]
    return out


def _values(env, refs):
    """Value references (``r >= 0``: state cell, ``r < 0``: constant) -> tensor ``[num_envs, len(refs)]``."""
    xp = env.xp
    refs = xp.as_tensor(refs, device=env.device).long().reshape(-1)
    if not hasattr(env, "_consts_dev"):
        env._consts_dev = env.engine._from_numpy(env.program.consts)
    state = env.engine.state[:, refs.clamp(min=0)]
    const = env._consts_dev[(-refs - 1).clamp(min=0)]
    return xp.where(refs >= 0, state, const.expand(env.num_envs, -1))


def get_polynomial_costs(env):
    """``objective.py:34-45``: ``[num_envs, 2 n_poly]`` = p-costs then q-costs of every ``poly_cost`` row."""
    sc, xp = env.program.scoring, env.xp
    n = sc["n_poly"]
    if n == 0:
        return xp.zeros((env.num_envs, 0), dtype=xp.float64, device=env.device)
    mul = lambda k: xp.as_tensor(sc[k], device=env.device)
    p = _values(env, sc["poly_p"]) * mul("poly_p_mul")
    q = _values(env, sc["poly_q"]) * mul("poly_q_mul")
    c = _values(env, sc["poly_coef"]).reshape(env.num_envs, n, 6)
    return xp.cat([c[..., 0] + c[..., 1] * p + c[..., 2] * p * p,
                   c[..., 3] + c[..., 4] * q + c[..., 5] * q * q], dim=1)


def get_piecewise_linear_costs(env):
    """``objective.py:57-77`` (including its missing sign test on the far side)."""
    sc, xp = env.program.scoring, env.xp
    n, n_seg = sc["n_pwl"], sc["n_pwl_seg"]
    if n == 0:
        return xp.zeros((env.num_envs, 0), dtype=xp.float64, device=env.device)
    v = _values(env, sc["pwl_v"]) * xp.as_tensor(sc["pwl_v_mul"], device=env.device)
    seg = _values(env, sc["pwl_seg"]).reshape(env.num_envs, n, n_seg, 3)
    sign, mag = xp.sign(v), v.abs()
    total = xp.zeros_like(v)
    for k in range(n_seg):
        lo, hi, price = seg[:, :, k, 0], seg[:, :, k, 1], seg[:, :, k, 2]
        near, far = xp.minimum(lo.abs(), hi.abs()), xp.maximum(lo.abs(), hi.abs())
        beyond = mag > far
        within = (mag > near) & (sign == xp.sign(lo + hi)) & ~beyond
        total = total + xp.where(beyond, sign * (hi - lo) * price, xp.zeros_like(v))
        total = total + xp.where(within, sign * (mag - near) * price, xp.zeros_like(v))
    return total


def get_pandapower_costs(env):
    """``objective.py:6-31`` for all environments: ``[num_envs, 2 n_poly + n_pwl]``; sum over dim 1 for
    the total costs (the env's objective is minus that, ``opf_env.py:493-500``)."""
    return env.xp.cat([get_polynomial_costs(env), get_piecewise_linear_costs(env)], dim=1)
