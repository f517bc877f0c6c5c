"""Several batched envs of different grids stepped as one batch (BASELINE config 4:
"MaxRenewable + QMarket mixed batch, fused constraint/objective/reward kernel").  Each member keeps
its own compiled grid and buffers.  ``step`` runs kernel 1 of ALL members as one launch, the members'
power flows (different kernels per grid size) on separate streams, and kernel 5 -- constraints,
objective, reward, observation -- of ALL members as one launch (``opfg_assemble_mixed`` /
``opfg_score_mixed``: every CTA picks its member's descriptor from a device table).  Outputs are
concatenated along the env axis (observations padded to the widest member); episode statistics add up.
``fused_launches=False`` steps the members independently on their streams instead."""
from __future__ import annotations

import ctypes as C

from . import capi
from .opf_env import BatchedOpfEnv


class MixedBatchEnv:
    def __init__(self, envs, fused_launches: bool = True):
        self.envs = list(envs)
        self.fused_launches = bool(fused_launches) and all(
            e._custom_solver is None and e._custom_objective is None and type(e).step is BatchedOpfEnv.step
            for e in self.envs)       # subclasses with their own step (multi-stage, N-1) are stepped member by member
        self._mixed = None
        self._host = None              # pinned host buffers of step_host, allocated on first use
        self.num_envs = sum(e.num_envs for e in self.envs)
        self.xp = self.envs[0].xp
        self.device = self.envs[0].device
        self.splits = [e.num_envs for e in self.envs]
        self.n_obs = max(e.single_observation_space.shape[0] for e in self.envs)
        self.n_act = max(e.single_action_space.shape[0] for e in self.envs)
        cuda = getattr(self.device, "type", "cpu") == "cuda"
        self._streams = [self.xp.cuda.Stream(device=self.device) for _ in self.envs] if cuda else None

    def _each(self, fn):
        """Run ``fn(i, env)`` for every member, each on its own stream; join on the caller's stream."""
        if self._streams is None:
            return [fn(i, e) for i, e in enumerate(self.envs)]
        xp = self.xp
        cur = xp.cuda.current_stream(self.device)
        start = xp.cuda.Event()
        start.record(cur)
        outs = []
        for i, (e, s) in enumerate(zip(self.envs, self._streams)):
            with xp.cuda.stream(s):
                s.wait_event(start)
                outs.append(fn(i, e))
                done = xp.cuda.Event()
                done.record(s)
            cur.wait_event(done)
        return outs

    def _pad(self, obs):
        xp = self.xp
        out = []
        for o in obs:
            if o.shape[1] < self.n_obs:
                pad = xp.full((o.shape[0], self.n_obs - o.shape[1]), float("nan"), dtype=o.dtype, device=o.device)
                o = xp.cat([o, pad], dim=1)
            out.append(o)
        return xp.cat(out, dim=0)

    def reset(self, seed=None, options=None):
        outs = self._each(lambda i, e: e.reset(seed=seed, options=options))
        return self._pad([o[0] for o in outs]), {}

    def _mixed_handle(self):
        if self._mixed is None:
            lib = self.envs[0].engine.lib
            n = len(self.envs)
            grids = (C.c_void_p * n)(*[e.engine.handle for e in self.envs])
            batches = (capi.Batch * n)(*[e.engine.batch_final for e in self.envs])
            h = C.c_void_p()
            capi.check(lib, lib.opfg_mixed_create(n, grids, batches, C.byref(h)))
            self._mixed = (lib, h)
        return self._mixed

    def _step_fused(self, parts):
        """kernel 1 of all members: one launch; power flows per member (own streams); kernel 5 of all
        members: one launch."""
        lib, h = self._mixed_handle()
        n = len(self.envs)
        acts = [e._step_begin(parts[i][:, :e.single_action_space.shape[0]]) for i, e in enumerate(self.envs)]
        batches = (capi.Batch * n)(*[e.engine.batch_final for e in self.envs])
        stream = self.envs[0].engine._stream()
        capi.check(lib, lib.opfg_assemble_mixed(h, batches, stream))
        self._each(lambda i, e: e.engine.pf_solve(e.engine.batch_final))
        capi.check(lib, lib.opfg_score_mixed(h, batches, stream))
        return [e._step_end(a) for e, a in zip(self.envs, acts)]

    def step(self, actions):
        """``actions[num_envs, n_act]``: member i reads its rows and its first n_act_i columns."""
        xp = self.xp
        act = xp.as_tensor(actions, device=self.device)
        parts = act.split(self.splits, dim=0)
        if self.fused_launches:
            outs = self._step_fused(parts)
        else:
            outs = self._each(lambda i, e: e.step(parts[i][:, :e.single_action_space.shape[0]]))
        obs = self._pad([o[0] for o in outs])
        reward, term, trunc = (xp.cat([o[k] for o in outs]) for k in (1, 2, 3))
        info = {"member": outs, "converged": xp.cat([o[4]["converged"] for o in outs]),
                "cost": xp.cat([o[4]["cost"] for o in outs])}
        return obs, reward, term, trunc, info

    def step_host(self, actions):
        """Host-buffer form (``BatchedOpfEnv.step_host``): the actions go to the device from pinned memory, the
        batch takes ONE ``step`` (kernel 1 / kernel 5 as single launches over all members), and observations
        (padded with NaN to the widest member), reward, flags and cost come back into pinned buffers that are
        allocated once -- one synchronisation per call, no host-side concatenation.  Returns numpy views of
        those buffers (valid until the next call)."""
        xp = self.xp
        cuda = getattr(self.device, "type", "cpu") == "cuda"
        act = xp.as_tensor(actions)
        if self._host is None:
            pin = dict(pin_memory=True) if cuda else {}
            self._host = dict(
                act=None, obs=None, reward=xp.empty(self.num_envs, dtype=xp.float64, **pin),
                term=xp.empty(self.num_envs, dtype=xp.bool, **pin), trunc=xp.empty(self.num_envs, dtype=xp.bool, **pin),
                cost=xp.empty(self.num_envs, dtype=xp.float64, **pin),
                conv=xp.empty(self.num_envs, dtype=xp.bool, **pin), pin=pin)
        h = self._host
        if not (cuda and act.is_pinned()):          # pageable input: through a pinned buffer of its own dtype
            if h["act"] is None or h["act"].shape != act.shape or h["act"].dtype != act.dtype:
                h["act"] = xp.empty(tuple(act.shape), dtype=act.dtype, **h["pin"])
            h["act"].copy_(act)
            act = h["act"]
        obs, reward, term, trunc, info = self.step(act.to(self.device, non_blocking=True).to(xp.float64))
        if h["obs"] is None or h["obs"].dtype != obs.dtype:
            h["obs"] = xp.empty(tuple(obs.shape), dtype=obs.dtype, **h["pin"])
        for dst, src in ((h["obs"], obs), (h["reward"], reward), (h["term"], term), (h["trunc"], trunc),
                         (h["cost"], info["cost"]), (h["conv"], info["converged"])):
            dst.copy_(src.to(dst.dtype) if src.dtype != dst.dtype else src, non_blocking=True)
        if cuda:
            xp.cuda.current_stream(self.device).synchronize()
        return (h["obs"].numpy(), h["reward"].numpy(), h["term"].numpy(), h["trunc"].numpy(),
                {"cost": h["cost"].numpy(), "converged": h["conv"].numpy()})

    def episode_statistics(self, reduce=True):
        stats = [e.episode_statistics(reduce=reduce) for e in self.envs]
        steps = sum(s["steps"] for s in stats)
        return {"steps": steps, "converged": sum(s["converged"] for s in stats),
                "valid": sum(s["valid"] for s in stats),
                "mean_reward": sum(s["mean_reward"] * s["converged"] for s in stats) /
                max(sum(s["converged"] for s in stats), 1.0),
                "members": stats}

    def close(self):
        if self._mixed is not None:
            self._mixed[0].opfg_mixed_destroy(self._mixed[1])
            self._mixed = None
        for e in self.envs:
            e.close()
