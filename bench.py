#!/usr/bin/env python
"""Headline benchmark: AC power-flow env steps/s of the batched VoltageControl
environment (BASELINE.json configs[1]: 32 768 envs of one 122-bus MV grid per GPU).

  python bench.py --gpus N --steps K --warmup W          # this repo's CUDA path
  python bench.py --impl reference ...                    # CPU arm: oracle port on host cores
  torchrun --nproc-per-node N bench.py --gpus N ...       # one rank per GPU, weak scaling

One "step" = `env.step(actions)` for the whole batch: apply actions, batched
Newton-Raphson power flow, fused scoring/observation, same-step auto-reset
(sample a new grid state for every env).  Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

# stdout carries exactly one JSON line: NCCL's own banner ("NCCL version ...") goes to stderr
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "ac_power_flow_env_steps_per_sec"
UNIT = "env_steps/s"
ENVS_PER_GPU = 32768
WORKLOAD = ("VoltageControl, synthetic stand-in of SimBench 1-MV-semiurb--1-sw (122 buses, 442 obs, "
            "14 actions), full_uniform sampling, U[0,1] actions")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--envs-per-gpu", type=int, default=ENVS_PER_GPU)
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------- CPU arm
def cpu_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def run_cpu_pool(n_per_worker: int, rounds: int, warm_rounds: int = 1):
    """Oracle env port under multiprocessing.Pool(all cores); returns (steps/s, total steps,
    converged, seconds, cores, per-round seconds)."""
    import multiprocessing as mp
    from oracle import env_port
    cores = cpu_cores()
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        for _ in range(warm_rounds):
            pool.map(env_port._worker_steps, [(s, 2) for s in range(cores)])
        times, done, ok = [], 0, 0
        for _ in range(rounds):
            t0 = time.perf_counter()
            out = pool.map(env_port._worker_steps, [(s, n_per_worker) for s in range(cores)])
            times.append(time.perf_counter() - t0)
            done += sum(o[0] for o in out)
            ok += sum(o[1] for o in out)
    total = sum(times)
    return done / total, done, ok, total, cores, times


def reference_arm(args):
    """--impl reference: the reference's CPU path for the same metric.  pandapower is not
    installable on this image, so the arm is the oracle port (kind "port")."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = cpu_cores()
    per_worker = 8
    rate, done, ok, secs, cores, times = run_cpu_pool(per_worker, rounds=args.steps,
                                                      warm_rounds=max(1, min(args.warmup, 2)))
    sample = (f"{args.steps} bench steps x {cores} workers x {per_worker} (reset+step) of the "
              f"oracle env port = {done} env steps in {secs:.1f} s")
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": WORKLOAD, "envs_per_step": cores * per_worker},
            "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "converged_share": ok / max(done, 1)}
    print(json.dumps(line))


# ------------------------------------------------------------------------ clocks
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6),
                              ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# -------------------------------------------------------------------------- GPU arm
def gpu_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    affinity = "default"
    if world > 1 and not os.environ.get("OPFG_NO_NUMA_BIND"):
        # One process per GPU: keep it (and the pinned buffers it is about to allocate, first touch)
        # on the CPU socket the GPU hangs off, so the 58 MB/step of observations of eight ranks do not
        # cross the socket interconnect.
        try:
            import pynvml
            pynvml.nvmlInit()
            pr = torch.cuda.get_device_properties(local_rank)
            bus_id = f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
            pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByPciBusId(bus_id.encode()))
            affinity = f"nvml numa-local ({len(os.sched_getaffinity(0))} cpus)"
        except Exception as exc:              # affinity is an optimisation, never a requirement
            affinity = f"default ({type(exc).__name__})"
    if world > 1:
        # NCCL prints its version banner to stdout while the communicator is created (it honours
        # NCCL_DEBUG_FILE only above the VERSION level): stdout points at stderr for that moment
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    dev = torch.device("cuda", local_rank)

    import __graft_entry__
    if not os.path.exists(__graft_entry__.LIB):
        if rank == 0:
            __graft_entry__.build()
        if world > 1:
            dist.barrier()
    from opfgym_b200 import envs

    B = args.envs_per_gpu
    env = envs.VoltageControl(num_envs=B, train_data="full_uniform", test_data="full_uniform",
                              n_profile_steps=672, rank=rank, world_size=world, device=dev,
                              seed=1234, copy_outputs=False)
    eng = env.engine
    n_act, n_obs = env.single_action_space.shape[0], env.single_observation_space.shape[0]
    # action pool, resident in HBM before the timed region (Philox, stream = global env id)
    pool = [torch.empty(B, n_act, dtype=torch.float64, device=dev) for _ in range(8)]
    for i, a in enumerate(pool):
        eng.philox_uniform(a, 4321, env.first_env, 1000 + i)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    env.reset(seed=1234)

    # ---- device-resident throughput ("value") ------------------------------------------
    def dev_step(i):
        env.step(pool[i % len(pool)])

    for i in range(args.warmup):
        dev_step(i)
    env.reset_statistics()
    eng.pf_events = []
    launches0 = eng.launch_count()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    ms_total = timed(dev_step, args.steps)
    clock_info = clocks.stop() if rank == 0 else None
    launches = eng.launch_count() - launches0
    pf_ms = sum(a.elapsed_time(b) for a, b in eng.pf_events) / max(len(eng.pf_events), 1)
    eng.pf_events = None
    stats = env.episode_statistics(reduce=True)
    value = world * B * args.steps / (ms_total * 1e-3)

    # ---- end to end through the public API with HOST buffers ("e2e") ---------------------
    # env.step_host: actions from a pinned host buffer, results (next observation, reward, cost,
    # converged) back in pinned host memory, stream-synchronised before it returns -- every step
    # float32 actions: the dtype of the reference's action space (gym.spaces.Box default, opf_env.py:130)
    h_act = [torch.rand(B, n_act, dtype=torch.float32).pin_memory() for _ in range(4)]
    e2e_sink = []

    def e2e_step(i):
        obs, reward, term, trunc, info = env.step_host(h_act[i % len(h_act)])
        e2e_sink[:] = [obs, reward, info["converged"]]   # numpy views of the pinned result buffers

    for i in range(min(args.warmup, 3)):
        e2e_step(i)
    e2e_steps = max(5, args.steps // 2)
    ms_e2e = timed(e2e_step, e2e_steps)
    e2e_value = world * B * e2e_steps / (ms_e2e * 1e-3)
    h2d = B * n_act * 4
    d2h = B * n_obs * 4 + B * 8 + B * 8 + B      # observation f32, reward, cost, converged

    # ---- FP64 peak probe (roofline denominator not in MEASURED_PEAKS.json) -----------------
    fp64_tflops = None
    if rank == 0:
        blocks, iters = 148 * 16, 20000
        eng.fp64_probe(blocks, 200)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.fp64_probe(blocks, iters)
        e1.record()
        torch.cuda.synchronize()
        fp64_tflops = 2.0 * 8 * iters * blocks * 256 / (e0.elapsed_time(e1) * 1e-3) / 1e12

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    info = eng.info
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        hbm_peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    else:
        hbm_peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    bytes_per_step = info["bytes_per_step"]
    achieved_gbs = bytes_per_step * B / (pf_ms * 1e-3) / 1e9
    iters = stats["mean_iterations"]
    flops_per_step = info["flops_per_iter"] * (iters + 1) + info["flops_score"]
    roofline = {"bound": "hbm", "kernel": "opfg_pf_solve = k_dc_start (DC-start GEMM, ~5 %) + k_pf_multi (fused mismatch + Jacobian + block sparse LU + solves)",
                "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": achieved_gbs / hbm_peak,
                "peak_source": peak_src, "traffic": 84.16e6 * B / 32768,
                "traffic_source": "profiles/r01i_k_pf_multi_metrics.csv (dram read 64.12 + write 20.04 MB per launch of 32768 envs)",
                "kernel_ms": pf_ms, "kernel_share_of_step": pf_ms / (ms_total / args.steps),
                "algorithmic_bytes_per_env_step": bytes_per_step,
                "note": "latency-bound FP64 sparse path: HBM fraction is small by construction "
                        "(SURVEY.md 8d); see fp64 for the pipe-side figure"}
    fp64 = {"flops_per_env_step": flops_per_step, "mean_nr_iterations": iters,
            "achieved_tflops_step": flops_per_step * value / world / 1e12,
            "achieved_tflops_kernel": info["flops_per_iter"] * (iters + 1) * B / (pf_ms * 1e-3) / 1e12,
            "peak_tflops_measured": fp64_tflops,
            "frac_kernel": (info["flops_per_iter"] * (iters + 1) * B / (pf_ms * 1e-3) / 1e12) / fp64_tflops}

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        per_worker = 8
        rate1, _, _, secs1, cores, _ = run_cpu_pool(per_worker, rounds=1)
        rounds = max(1, int(args.cpu_seconds / max(secs1, 1e-3)))
        rate, done, ok, secs, cores, _ = run_cpu_pool(per_worker, rounds=rounds, warm_rounds=0)
        cpu_baseline = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"{done} env steps (reset+step of the oracle env port, same grid and "
                                  f"sampler) on {cores} host cores in {secs:.1f} s",
                        "converged_share": ok / max(done, 1)}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "envs_per_gpu": B, "global_batch": world * B,
                       "parallelism": f"env-sharded x{world}, no data-path collective", "cpu_affinity": affinity,
                       "l2": f"inputs larger than L2: per-GPU state matrix "
                             f"{B * info['n_state'] * 8 / 1e6:.0f} MB is re-sampled every step",
                       "n_state": info["n_state"], "n_levels": info["n_levels"],
                       "n_blocks": info["n_blocks"], "threads_per_env": info["threads_per_env"],
                       "smem_bytes_pf": info["smem_bytes_pf"]},
            "clocks": clock_info,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / e2e_steps},
            "gpu_launches": launches,
            "roofline": roofline, "fp64": fp64, "cpu_baseline": cpu_baseline,
            "converged_share": stats["converged_share"], "valid_share": stats["valid_share"],
            "mean_reward": stats["mean_reward"]}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        reference_arm(args)
    else:
        gpu_arm(args)


if __name__ == "__main__":
    main()
