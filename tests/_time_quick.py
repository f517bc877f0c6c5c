"""Quick kernel timing of engine variants (developer tool, run under gpurun)."""
import sys; sys.path.insert(0, '.')
import torch
from tests import common
from opfgym_b200.engine import Engine

def fill(case, eng, B):
    for t, c in common.SAMPLED:
        df = case.net[t]
        if len(df):
            lo = torch.tensor(df["min_min_" + c].to_numpy() / df.scaling.to_numpy(), device="cuda")
            hi = torch.tensor(df["max_max_" + c].to_numpy() / df.scaling.to_numpy(), device="cuda")
            eng.column(t, c).copy_(lo + (hi - lo) * torch.rand(B, len(df), device="cuda", dtype=torch.float64))
    eng.actions.uniform_(0, 1)

def timeit(fn, n=20, w=5):
    for _ in range(w): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "1-MV-semiurb--1-sw"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
    threads = [int(x) for x in sys.argv[3].split(',')] if len(sys.argv) > 3 else [32, 64, 128]
    orders = [int(x) for x in sys.argv[4].split(',')] if len(sys.argv) > 4 else [0]
    case = common.make_case(name)
    for o in orders:
      for t in threads:
        v = dict(ordering=o, threads_per_env=t)
        try:
            eng = Engine(case.program, B, **v)
        except Exception as e:
            print(v, 'ERR', e); continue
        fill(case, eng, B)
        eng.assemble()
        ms = timeit(eng.pf_solve)
        i = eng.info
        ms2 = timeit(eng.score)
        print(f"{name} {v} levels={i['n_levels']} blocks={i['n_blocks']} smem={i['smem_bytes_pf']} "
              f"pf={ms:.3f} ms -> {B/ms*1e3:.3e} env/s  iters={eng.iterations.float().mean().item():.2f}  score={ms2:.3f} ms", flush=True)
        eng.close(); del eng


if __name__ == "__main__":
    main()
