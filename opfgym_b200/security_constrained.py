"""N-1 security-constrained scoring on the batched engine.

Reference ``opfgym/security_constrained.py:37-68``: after the base-case
violations, every listed element is taken out of service in turn, the power flow
is re-run and the violations are accumulated -- valids AND-ed, violations and
penalties summed; a failed contingency power flow zeroes the valids and adds
``not_converged_penalty`` to both sums.  Here one contingency is ONE more pass
of kernels 1/2-4/5 over the whole batch with the element's ``in_service`` cell
cleared (per-environment branch parameters, see ``opfg_set_dynamic_branches``);
the base case and all contingencies reuse the same compiled grid.

An outage that islands part of the grid (every line outage on a radial feeder does) is handled the way
pandapower handles it: kernel 1 finds the buses cut off from every slack bus per environment, the power
flow solves the rest, the dropped buses and their branches report NaN and violate nothing
(``tests/test_islands.py``).
"""
from __future__ import annotations

import numpy as np

from .opf_env import BatchedOpfEnv


class SecurityConstrainedBatchedOpfEnv(BatchedOpfEnv):
    def __init__(self, *args, n_minus_one_keys, not_converged_penalty: float = 1, **kwargs):
        self.n_minus_one_keys = [(t, c, np.asarray(i)) for t, c, i in n_minus_one_keys]
        for unit_type, column, _ in self.n_minus_one_keys:
            if (unit_type, column) not in (("line", "in_service"), ("trafo", "in_service"), ("switch", "closed")):
                raise NotImplementedError("contingencies are line/trafo in_service cells or closed cells of "
                                          "line-bus / trafo-bus switches")
        self.not_converged_penalty = float(not_converged_penalty)
        dyn = list(kwargs.pop("dynamic_columns", ()))
        dyn += [(t, c) for t, c, _ in self.n_minus_one_keys if (t, c) not in dyn]
        kwargs.setdefault("prefetch_reset", False)   # contingency passes reuse the state buffer
        super().__init__(*args, dynamic_columns=dyn, **kwargs)

    def step(self, actions):
        obs, reward, terminated, truncated, info = self._step_with_contingencies(actions)
        return obs, reward, terminated, truncated, info

    def step_host(self, actions=None):
        raise NotImplementedError("step_host would skip the contingency passes; use step()")

    def _step_with_contingencies(self, actions):
        xp, e = self.xp, self.engine
        act = xp.as_tensor(actions, device=self.device)
        e.actions.copy_(act.reshape(e.actions.shape))
        e.step(final_obs=True)
        self.power_flow_available = True
        self._results = e
        aux = e.enable_aux_results()      # contingency passes: scratch results, no statistics
        nc = max(len(self.constraints), 1)
        base_ok = e.converged.bool().clone()
        objective = e.objective.clone()
        valids = e.valids[:, :nc].bool().clone()
        violations = e.violations[:, :nc].clone()
        penalties = e.penalties[:, :nc].clone()
        final_obs = self._obs_out(final=True).clone()
        for unit_type, column, idxs in self.n_minus_one_keys:
            cells = self.col(unit_type, column)
            for pos in self.positions(unit_type, idxs):
                was_on = cells[:, pos] != 0                      # security_constrained.py:46-48
                saved = cells[:, pos].clone()
                cells[:, pos] = 0.0
                e.assemble(apply_actions=False)
                e.pf_solve(e.batch_aux)
                e.score(e.batch_aux)
                ok = aux.converged.bool()
                use = was_on & ok
                fail = was_on & ~ok
                valids = xp.where(use[:, None], valids & aux.valids[:, :nc].bool(), valids)
                violations = violations + xp.where(use[:, None], aux.violations[:, :nc], 0.0)
                penalties = penalties + xp.where(use[:, None], aux.penalties[:, :nc], 0.0)
                valids = valids & ~fail[:, None]                 # :59-64
                violations = violations + fail[:, None] * self.not_converged_penalty
                penalties = penalties + fail[:, None] * self.not_converged_penalty
                cells[:, pos] = saved
        penalty = penalties.sum(dim=1)
        valid = valids.all(dim=1)
        reward = self.reward_function.batched(objective, penalty, valid)
        cost = self.reward_function.batched_cost(penalty, valid)
        nan = xp.full_like(reward, float("nan"))
        reward = xp.where(base_ok, reward, nan)
        info = {"valids": valids, "violations": violations, "unscaled_penalties": penalties,
                "cost": cost, "converged": base_ok, "final_obs": final_obs}
        terminated = xp.ones(self.num_envs, dtype=xp.bool, device=self.device)
        truncated = xp.zeros(self.num_envs, dtype=xp.bool, device=self.device)
        self._begin_episode()
        return self._obs_out(), reward, terminated, truncated, info
