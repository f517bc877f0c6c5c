"""Objective of the batched engine.

The reference's default objective ``get_pandapower_costs(net)``
(``opfgym/objective.py:6-31``) is a function of the net's ``poly_cost`` /
``pwl_cost`` tables.  Here those two tables ARE the objective description: the
compiler (``opfgym_b200.compiler``) flattens them into device tables and kernel
5 evaluates ``[p-costs, q-costs, pwl-costs]`` per
environment.  ``cost_vector_layout`` tells the order of the reference's cost
vector for a given net.
"""


def cost_vector_layout(net) -> list[tuple[str, int, str]]:
    """``[(et, element, 'p'|'q'|'pwl'), ...]`` in the order of the reference's
    ``get_pandapower_costs`` output."""
    pc, pw = net.poly_cost, net.pwl_cost
    out = [(et, int(el), "p") for et, el in zip(pc.et, pc.element)]
    out += [(et, int(el), "q") for et, el in zip(pc.et, pc.element)]
    out += [(et, int(el), "pwl") for et, el in zip(pw.et, pw.element)]
    return out
