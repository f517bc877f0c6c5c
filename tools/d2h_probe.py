#!/usr/bin/env python
"""Host-ingest ceiling of the box: pinned device->host copies of one observation matrix per rank.

  torchrun --nproc-per-node N tools/d2h_probe.py [MB]

Per rank: a device buffer of MB megabytes (default 58 = 32 768 envs x 442 f32 observations) copied to
pinned host memory 40 times, (a) every rank in turn while the others idle, (b) all ranks at once,
(c) all at once in two halves on two streams, (d) all at once into write-combined pinned memory
(cudaHostAllocWriteCombined), (e) all at once with the rank bound to the GPU's NUMA-local cores (NVML).
Prints GB/s per rank and in total: what `bench.py`'s e2e number at N ranks can reach at most."""
import ctypes, os, sys, time
import torch
import torch.distributed as dist

mb = float(sys.argv[1]) if len(sys.argv) > 1 else 58.0
world, rank, lr = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n = int(mb * 1e6) // 4
dev = torch.empty(n, dtype=torch.float32, device="cuda").normal_()
REP = 40


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def gbs(fn):
    for _ in range(3):
        fn()
    barrier()
    t0 = time.perf_counter()
    for _ in range(REP):
        fn()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return REP * n * 4 / dt / 1e9


def gather(x):
    t = torch.tensor([x], device="cuda", dtype=torch.float64)
    if world > 1:
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [float(o.item()) for o in out]
    return [x]


def report(name, mine):
    vals = gather(mine)
    if rank == 0:
        print(f"{name:34s} per rank {' '.join(f'{v:5.1f}' for v in vals)}  total {sum(vals):6.1f} GB/s", flush=True)


host = torch.empty(n, dtype=torch.float32).pin_memory()
plain = lambda: (host.copy_(dev, non_blocking=True), torch.cuda.synchronize())
# (a) one rank at a time
alone = 0.0
for r in range(world):
    if world > 1:
        dist.barrier()
    if r == rank:
        for _ in range(3):
            plain()
        t0 = time.perf_counter()
        for _ in range(REP):
            plain()
        alone = REP * n * 4 / (time.perf_counter() - t0) / 1e9
    if world > 1:
        dist.barrier()
report("(a) alone, one rank at a time", alone)
report("(b) all ranks at once", gbs(plain))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
half = n // 2


def two_streams():
    with torch.cuda.stream(s1):
        host[:half].copy_(dev[:half], non_blocking=True)
    with torch.cuda.stream(s2):
        host[half:].copy_(dev[half:], non_blocking=True)
    s1.synchronize(); s2.synchronize()


report("(c) all at once, two streams", gbs(two_streams))
try:
    rt = ctypes.CDLL("libcudart.so.12")
    ptr = ctypes.c_void_p()
    assert rt.cudaHostAlloc(ctypes.byref(ptr), ctypes.c_size_t(n * 4), ctypes.c_uint(4)) == 0   # write-combined
    dptr = ctypes.c_void_p(dev.data_ptr())
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    rt.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
    wc = lambda: (rt.cudaMemcpyAsync(ptr, dptr, n * 4, 2, stream), torch.cuda.synchronize())
    report("(d) all at once, write-combined", gbs(wc))
except Exception as exc:     # noqa: BLE001
    if rank == 0:
        print("(d) write-combined: unavailable", type(exc).__name__, exc)
try:
    import pynvml
    pynvml.nvmlInit()
    pr = torch.cuda.get_device_properties(lr)
    bus_id = f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
    pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByPciBusId(bus_id.encode()))
    host2 = torch.empty(n, dtype=torch.float32).pin_memory()        # first touch on the local node
    local = lambda: (host2.copy_(dev, non_blocking=True), torch.cuda.synchronize())
    report(f"(e) all at once, NUMA-local ({len(os.sched_getaffinity(0))} cpus)", gbs(local))
except Exception as exc:     # noqa: BLE001
    if rank == 0:
        print("(e) NUMA binding: unavailable", type(exc).__name__, exc)
if rank == 0:
    try:
        print(open("/sys/devices/system/node/online").read().strip(), "NUMA nodes online;", os.cpu_count(), "cpus")
    except OSError:
        pass
if world > 1:
    dist.destroy_process_group()
