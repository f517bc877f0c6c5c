"""Torch-facing handle on one compiled grid + one batch of environment buffers.

PyTorch is plumbing here: it owns the device memory and the stream.  Every
numeric step is a launch of a hand-written kernel in ``libopfg_b200.so`` through
the C ABI (``include/opfg_b200.h``).  There is no CPU path: constructing an
``Engine`` without a CUDA device, or without the built library, raises.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .compiler import EnvProgram, fill_descs


class ResetPlan:
    def __init__(self, lib, handle, keep):
        self.lib, self.handle, self.keep = lib, handle, keep     # `keep`: device arrays the plan points to

    def __del__(self):
        try:
            self.lib.opfg_reset_plan_destroy(self.handle)
        except Exception:
            pass


def check_engine_class(engine_cls):
    """``engine_cls=`` is a TEST seam (the CPU tier builds the same kernel sources with g++ to check schedules
    and marshalling without a GPU).  The product has one engine: anything else must mark itself as test
    infrastructure, so that a CPU engine cannot slip into a product path unnoticed."""
    if engine_cls is not Engine and not getattr(engine_cls, "OPFG_TEST_ENGINE", False):
        raise TypeError("engine_cls is a test seam: only opfgym_b200.engine.Engine (CUDA) or a class that sets "
                        "OPFG_TEST_ENGINE = True (tests/hostsim) is accepted")


class Engine:
    STATS_SLOTS = 128
    def __init__(self, program: EnvProgram, num_envs: int, device=None,
                 tolerance_mva: float = 1e-8, max_iteration: int = 10, init: str = "dc",
                 enforce_q_lims: bool = True, threads_per_env: int = 0, ordering: int = 0,
                 pf_kernel: str | int = 0,
                 obs_dtype: str = "float32", lib=None):
        self.program = program
        self.num_envs = int(num_envs)
        self.lib = lib if lib is not None else capi.load()
        self._setup_device(device)
        gd, ad, sd, dd, keep = fill_descs(
            capi, program, tol_pu=tolerance_mva / program.ppc.base_mva,
            max_iter=max_iteration, init_dc=(init == "dc"), enforce_q_lims=enforce_q_lims,
            threads_per_env=threads_per_env, ordering=ordering,
            pf_kernel={"auto": 0, "cta": 1, "lanes": 2, "radial": 3}.get(pf_kernel, pf_kernel))
        handle = C.c_void_p()
        capi.check(self.lib, self.lib.opfg_grid_create(C.byref(gd), C.byref(handle)))
        self.handle = handle
        capi.check(self.lib, self.lib.opfg_set_assembly(handle, C.byref(ad)))
        if dd is not None:
            capi.check(self.lib, self.lib.opfg_set_dynamic_branches(handle, C.byref(dd)))
        capi.check(self.lib, self.lib.opfg_set_scoring(handle, C.byref(sd)))
        del keep
        info = capi.GridInfo()
        capi.check(self.lib, self.lib.opfg_grid_info(handle, C.byref(info)))
        self.info = {name: getattr(info, name) for name, _ in capi.GridInfo._fields_}

        B, nb = self.num_envs, program.ppc.bus.shape[0]
        nc, n_state = len(program.constraints), program.layout.n
        self.obs_dtype = obs_dtype
        self.actions = self._zeros((B, max(program.n_act, 1)), "float64")
        self.state = self._from_numpy(np.tile(program.initial_state, (B, 1)))
        self.sbus = self._zeros((B, nb, 2), "float64")
        self.vm = self._zeros((B, nb), "float64")
        self.va = self._zeros((B, nb), "float64")
        self.converged = self._zeros((B,), "uint8")
        self.iterations = self._zeros((B,), "int32")
        self.reward = self._zeros((B,), "float64")
        self.objective = self._zeros((B,), "float64")
        self.penalty = self._zeros((B,), "float64")
        self.cost = self._zeros((B,), "float64")
        self.valids = self._zeros((B, max(nc, 1)), "uint8")
        self.violations = self._zeros((B, max(nc, 1)), "float64")
        self.penalties = self._zeros((B, max(nc, 1)), "float64")
        self.obs = self._zeros((B, max(program.n_obs, 1)), obs_dtype)
        # observation of the episode that just ended (kernel 5) vs. of the freshly reset one
        self.obs_final = self._zeros((B, max(program.n_obs, 1)), obs_dtype)
        # per-environment statistics are added to row (env % STATS_SLOTS): atomics on one row serialise
        self.stats = self._zeros((self.STATS_SLOTS, capi.N_STATS), "float64")
        self.n_constraints = nc
        self.objective_offset = None     # set by enable_objective_offset() (diff_objective)
        self.yval = self.bry = None
        if dd is not None:     # per-environment Ybus values / branch admittances (kernel 1 output)
            self.yval = self._zeros((B, self.info["nnz_y"], 2), "float64")
            self.bry = self._zeros((B, dd.n_dyn, 8), "float64")
        self.batch = capi.Batch(
            n_env=B, actions=self._ptr(self.actions), state=self._ptr(self.state),
            sbus=self._ptr(self.sbus), vm=self._ptr(self.vm), va=self._ptr(self.va),
            converged=self._ptr(self.converged), iterations=self._ptr(self.iterations),
            reward=self._ptr(self.reward), objective=self._ptr(self.objective),
            penalty=self._ptr(self.penalty), cost=self._ptr(self.cost),
            valids=self._ptr(self.valids), violations=self._ptr(self.violations),
            penalties=self._ptr(self.penalties),
            obs_f32=self._ptr(self.obs) if obs_dtype == "float32" else None,
            obs_f64=self._ptr(self.obs) if obs_dtype == "float64" else None,
            stats=self._ptr(self.stats), stats_slots=self.STATS_SLOTS,
            yval=self._ptr(self.yval) if self.yval is not None else None,
            bry=self._ptr(self.bry) if self.bry is not None else None)
        self.batch_final = capi.Batch.from_buffer_copy(self.batch)
        if obs_dtype == "float32":
            self.batch_final.obs_f32 = self._ptr(self.obs_final)
        else:
            self.batch_final.obs_f64 = self._ptr(self.obs_final)
        # reset applies its own (centre / random) action: separate buffer so that a reset prefetched on a
        # side stream never touches the agent's actions
        self.actions_reset = self._zeros((B, max(program.n_act, 1)), "float64")
        self.batch_setpoints = capi.Batch.from_buffer_copy(self.batch)
        self.batch_setpoints.sbus = None
        self.batch_setpoints.actions = self._ptr(self.actions_reset)
        self.batch_setpoints.absolute_actions = 1
        self._linked = [self.batch, self.batch_final, self.batch_setpoints]   # batches that follow select()
        self._states = [self.state]
        self._obs_bufs = [self.obs]        # every episode buffer has its own (reset) observation
        self.cur = 0
        self.aux = None
        self.batch_aux = None
        self.batch_nostats = None

    def enable_objective_offset(self):
        """[B] buffer subtracted from the objective in kernel 5 (``diff_objective``)."""
        if self.objective_offset is None:
            self.objective_offset = self._zeros((self.num_envs,), "float64")
            for b in self._linked:
                b.objective_offset = self._ptr(self.objective_offset)
        return self.objective_offset

    class _Aux:
        pass

    def enable_aux_results(self):
        """Second set of per-step result buffers WITHOUT statistics, for power flows that are not
        the agent's step: the reset power flow (``pf_for_obs``), contingency passes.  The buffers the
        last ``step`` filled (and may have handed out as aliases) stay untouched, and the episode
        statistics only count agent steps (round-1 advisor finding)."""
        if self.aux is None:
            a = self._Aux()
            for name in ("converged", "iterations", "reward", "objective", "penalty", "cost",
                         "valids", "violations", "penalties"):
                setattr(a, name, self._zeros(tuple(getattr(self, name).shape),
                                             str(getattr(self, name).dtype).replace("torch.", "")))
            self.aux = a
            self.batch_aux = capi.Batch.from_buffer_copy(self.batch)
            for name in ("converged", "iterations", "reward", "objective", "penalty", "cost",
                         "valids", "violations", "penalties"):
                setattr(self.batch_aux, name, self._ptr(getattr(a, name)))
            self.batch_aux.stats = None
            self._linked.append(self.batch_aux)
        return self.aux

    def nostats_batch(self):
        """The main batch without the statistics epilogue (``run_power_flow`` re-scores the current
        state; that is not an agent step)."""
        if self.batch_nostats is None:
            self.batch_nostats = capi.Batch.from_buffer_copy(self.batch)
            self.batch_nostats.stats = None
            self._linked.append(self.batch_nostats)
        return self.batch_nostats

    def enable_double_buffer(self, n: int = 2):
        """More state matrices: later episodes can be sampled while the current one is solved."""
        while len(self._states) < n:
            self._states.append(self.state.clone())
            self._obs_bufs.append(self.obs.clone())

    def select(self, index: int):
        """Make state buffer ``index`` the one every launch (and ``column()``) refers to."""
        self.cur = index
        self.state = self._states[index]
        self.obs = self._obs_bufs[index]
        ptr, obs_ptr = self._ptr(self.state), self._ptr(self.obs)
        f32 = str(self.obs.dtype).endswith("float32")
        for b in self._linked:
            b.state = ptr
            if b is not self.batch_final:          # that one writes the finished episode's observation
                if f32:
                    b.obs_f32 = obs_ptr
                else:
                    b.obs_f64 = obs_ptr

    # ------------------------------------------------------- device plumbing (torch)
    def _setup_device(self, device):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("opfgym_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.torch = torch
        self.device = torch.device(device if device is not None else "cuda:0")
        torch.cuda.set_device(self.device)

    def _zeros(self, shape, dtype):
        return self.torch.zeros(shape, dtype=getattr(self.torch, dtype), device=self.device)

    def _from_numpy(self, a):
        return self.torch.from_numpy(np.array(a, copy=True, order="C")).to(self.device)

    def _ptr(self, t):
        return C.c_void_p(t.data_ptr())

    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    # ------------------------------------------------------------------- launches
    def column(self, table: str, column: str):
        """View ``S[:, slice]`` of one (table, column): shape [B, n_rows]."""
        return self.state[:, self.program.layout.slice(table, column)]

    trace = None     # a list while the env layer records one episode reset for ``make_reset_plan``

    def make_reset_plan(self, trace, stream_base: int):
        """Fused-reset plan from a recorded sequence of sampler / row-program calls."""
        stages, keep = [], []
        for item in trace:
            if item[0] == "sample":
                _, slots, lo, hi, div, stream_id = item
                keep += [slots, lo, hi, div]
                stages.append(capi.ResetStage(kind=0, n_cols=int(slots.shape[0]), slots=self._ptr(slots).value,
                                              lo=self._ptr(lo).value, hi=self._ptr(hi).value,
                                              div=self._ptr(div).value,
                                              stream_offset=stream_id - stream_base))
            else:
                keep.append(item[1])
                stages.append(capi.ResetStage(kind=1, program=item[1].handle.value))
        arr = (capi.ResetStage * max(len(stages), 1))(*stages)
        h = C.c_void_p()
        capi.check(self.lib, self.lib.opfg_reset_plan_create(arr, len(stages), C.byref(h)))
        return ResetPlan(self.lib, h, keep)

    def reset_episode(self, plan, seed: int, first_env: int, stream_base: int, random_action: bool,
                      action_stream_offset: int):
        """ONE launch: sampler stages + hook programs + initial action + set-points + observation."""
        capi.check(self.lib, self.lib.opfg_reset_episode(
            self.handle, C.byref(self.batch_setpoints), plan.handle, seed, first_env, stream_base,
            int(random_action), action_stream_offset, self._stream()))

    def sample_uniform(self, slots, lo, hi, div, seed: int, first_env: int, stream_id: int, obs_pos=None):
        """``obs_pos`` (device int32, per sampled column: its position in the observation or -1): the sampled
        values are written into ``self.obs`` on the way (opfg_sample_uniform_obs)."""
        if self.trace is not None:
            self.trace.append(("sample", slots, lo, hi, div, stream_id))
        n = int(slots.shape[0])
        if obs_pos is None:
            capi.check(self.lib, self.lib.opfg_sample_uniform(
                seed, first_env, stream_id, self.num_envs, n, self._ptr(slots), self._ptr(lo),
                self._ptr(hi), self._ptr(div), self._ptr(self.state), self.program.layout.n,
                self._stream()))
            return
        f32 = str(self.obs.dtype).endswith("float32")
        capi.check(self.lib, self.lib.opfg_sample_uniform_obs(
            seed, first_env, stream_id, self.num_envs, n, self._ptr(slots), self._ptr(lo),
            self._ptr(hi), self._ptr(div), self._ptr(self.state), self.program.layout.n,
            self._ptr(obs_pos), self._ptr(self.obs) if f32 else None, None if f32 else self._ptr(self.obs),
            int(self.obs.shape[1]), self._stream()))

    def sample_profiles(self, slots, table, step, interp_r, pmin, pmax, noise_factor: float, noise_kind: int,
                        seed: int, first_env: int, stream_id: int):
        """ONE launch per profile table: row gather, interpolation, noise, clip (opf_env.py:317-372)."""
        capi.check(self.lib, self.lib.opfg_sample_profiles(
            seed, first_env, stream_id, self.num_envs, int(slots.shape[0]), self._ptr(slots), self._ptr(table),
            int(table.shape[0]), self._ptr(step), self._ptr(interp_r) if interp_r is not None else None,
            self._ptr(pmin), self._ptr(pmax), float(noise_factor), int(noise_kind), self._ptr(self.state),
            self.program.layout.n, self._stream()))

    def philox_uniform(self, out, seed: int, first_env: int, stream_id: int):
        capi.check(self.lib, self.lib.opfg_philox_uniform(
            seed, first_env, stream_id, out.shape[0], out.shape[1], self._ptr(out), self._stream()))

    def assemble(self, apply_actions: bool = True, scatter_sbus: bool = True, absolute: bool = False):
        """Kernel 1.  ``apply_actions=False`` only re-scatters Sbus from the current cells;
        ``scatter_sbus=False`` only writes the set-points (from ``actions_reset``, always as
        absolute set-points); ``absolute`` ignores a configured incremental step size."""
        batch = self.batch
        if not apply_actions:
            batch = capi.Batch.from_buffer_copy(self.batch)
            batch.actions = None
        elif not scatter_sbus:
            batch = self.batch_setpoints
        elif absolute:
            batch = capi.Batch.from_buffer_copy(self.batch)
            batch.absolute_actions = 1
        capi.check(self.lib, self.lib.opfg_assemble(self.handle, C.byref(batch), self._stream()))

    def pf_solve(self, batch=None):
        """Kernels 2-4.  With ``self.pf_events`` set to a list, CUDA events bracketing the launch(es) on the
        launching stream are appended to it (bench.py's roofline)."""
        if self.pf_events is not None:
            e0 = self.torch.cuda.Event(enable_timing=True)
            e1 = self.torch.cuda.Event(enable_timing=True)
            e0.record()
        capi.check(self.lib, self.lib.opfg_pf_solve(self.handle, C.byref(batch or self.batch), self._stream()))
        if self.pf_events is not None:
            e1.record()
            self.pf_events.append((e0, e1))

    def score(self, batch=None):
        capi.check(self.lib, self.lib.opfg_score(self.handle, C.byref(batch or self.batch), self._stream()))

    def observe(self):
        capi.check(self.lib, self.lib.opfg_observe(self.handle, C.byref(self.batch), self._stream()))

    def step(self, final_obs: bool = False):
        """assemble -> pf_solve -> score on the current stream (3 launches, no sync).
        ``final_obs``: kernel 5 writes its observation to ``obs_final`` (the env layer's
        auto-reset then fills ``obs`` without a copy).  With ``self.pf_events`` set to a list,
        CUDA events bracketing the power-flow kernel are appended to it (bench.py's roofline)."""
        batch = self.batch_final if final_obs else self.batch
        if self.pf_events is None:
            capi.check(self.lib, self.lib.opfg_step(self.handle, C.byref(batch), self._stream()))
            return
        self.assemble()
        self.pf_solve()
        capi.check(self.lib, self.lib.opfg_score(self.handle, C.byref(batch), self._stream()))

    pf_events = None

    def fp64_probe(self, n_blocks: int, iters: int):
        capi.check(self.lib, self.lib.opfg_fp64_probe(n_blocks, iters, self._ptr(self.stats), self._stream()))

    def symbolic(self):
        n, nl = self.info["n_nonref"], self.info["n_levels"]
        perm = np.zeros(n, np.int32)
        lvl = np.zeros(nl + 1, np.int32)
        capi.check(self.lib, self.lib.opfg_grid_symbolic(
            self.handle, perm.ctypes.data_as(C.POINTER(C.c_int32)),
            lvl.ctypes.data_as(C.POINTER(C.c_int32))))
        return perm, lvl

    def launch_count(self) -> int:
        return int(self.lib.opfg_launch_count())

    def close(self):
        if getattr(self, "handle", None):
            self.lib.opfg_grid_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
