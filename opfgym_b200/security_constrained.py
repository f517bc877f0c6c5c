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
    def __init__(self, *args, n_minus_one_keys, not_converged_penalty: float = 1,
                 batch_contingencies: bool | None = None, **kwargs):
        """``batch_contingencies``: solve ALL contingencies of all environments as one batch of
        ``num_envs x n_contingencies`` rows (three launches per step instead of three per contingency;
        SURVEY.md 8f rank 2: "the batch dimension becomes B x (1 + n_outage)").  ``None`` = automatic: on when
        that batch has at most 262 144 rows, else one pass per contingency over the ``num_envs`` rows."""
        self.batch_contingencies = batch_contingencies
        self._cont_engine = None
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

    def _rebuild_engine(self):
        super()._rebuild_engine()
        if self._cont_engine is not None:          # compiled for the previous program
            self._cont_engine.close()
            self._cont_engine = None

    def close(self):
        if self._cont_engine is not None:
            self._cont_engine.close()
            self._cont_engine = None
        super().close()

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
        cont_cells = [(self.program.layout.slice(t, c).start + int(pos))
                      for t, c, idxs in self.n_minus_one_keys for pos in self.positions(t, idxs)]
        batched = self.batch_contingencies
        if batched is None:
            batched = self.num_envs * len(cont_cells) <= 262144
        if batched and cont_cells:
            # every (environment, contingency) pair is one row of a second engine on the same compiled grid: the
            # row is the environment's state after the base case with the contingency's cell cleared
            B, K = self.num_envs, len(cont_cells)
            if self._cont_engine is None:
                self._cont_engine = self._engine_cls(self.program, B * K, **self._engine_args)
                self._cont_cells = xp.as_tensor(np.asarray(cont_cells), device=self.device)
                self._cont_k = xp.arange(K, device=self.device)
            big = self._cont_engine
            rows = big.state.view(B, K, -1)
            was_on = e.state[:, self._cont_cells] != 0           # security_constrained.py:46-48
            rows.copy_(e.state[:, None, :].expand(B, K, e.state.shape[1]))
            rows[:, self._cont_k, self._cont_cells] = 0.0
            big.assemble(apply_actions=False)
            big.pf_solve()
            big.score()
            ok = big.converged.view(B, K).bool()
            use = (was_on & ok)[:, :, None]
            fail = (was_on & ~ok)
            c_valid = big.valids.view(B, K, -1)[:, :, :nc].bool()
            valids = valids & (c_valid | ~use).all(dim=1) & ~fail.any(dim=1)[:, None]
            n_fail = fail.sum(dim=1, keepdim=True) * self.not_converged_penalty         # :59-64
            violations = violations + xp.where(use, big.violations.view(B, K, -1)[:, :, :nc], 0.0).sum(dim=1) + n_fail
            penalties = penalties + xp.where(use, big.penalties.view(B, K, -1)[:, :, :nc], 0.0).sum(dim=1) + n_fail
        for unit_type, column, idxs in (() if batched else self.n_minus_one_keys):
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
