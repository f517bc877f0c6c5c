#!/usr/bin/env python
"""Attribute an ncu SASS-level source page to CUDA source lines.

usage: ncu_lines.py <report.ncu-rep> <lib.so> <mangled-kernel-substring> [top_n]
Joins `ncu --page source --csv` (per-SASS-instruction counters, no line numbers in
CSV mode) with `nvdisasm -g` line info of the same cubin, by instruction order.
"""
import csv, io, os, re, subprocess, sys, tempfile, collections

rep, lib, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
lines = []
for f in os.listdir(tmp):
    if f.endswith(".cubin"):
        out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        cur, on, file_, line_ = None, False, "?", 0
        for l in out.splitlines():
            m = re.match(r"\.text\.(\S+):", l)
            if m:
                on = kname in m.group(1)
                continue
            if not on:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', l)
            if m:
                file_, line_ = os.path.basename(m.group(1)), int(m.group(2))
                continue
            m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
            if m:
                lines.append((file_, line_, m.group(2)))
extra = os.environ.get("NCU_FILTER", "").split()     # e.g. NCU_FILTER="-k regex:k_score" for multi-kernel reports
csvtxt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + extra, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(csvtxt)))
hdr_i = next(i for i, r in enumerate(rows) if "Source" in r and "Address" in r)
hdr = rows[hdr_i]
ci = {n: hdr.index(n) for n in ("Source", "# Samples", "Instructions Executed", "Thread Instructions Executed")}
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_")]
body = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
if len(body) != len(lines):
    print(f"warning: {len(body)} ncu rows vs {len(lines)} nvdisasm instructions", file=sys.stderr)
agg = collections.defaultdict(lambda: [0, 0, 0, collections.Counter(), 0, 0])
ti = ts = tw = 0
wf_i = hdr.index("L1 Wavefronts Shared") if "L1 Wavefronts Shared" in hdr else None
wfi_i = hdr.index("L1 Wavefronts Shared Ideal") if "L1 Wavefronts Shared Ideal" in hdr else None
sort_key = os.environ.get("NCU_SORT", "samples")      # samples | wavefronts | inst
for r, (f, ln, txt) in zip(body, lines):
    n = int(r[ci["Instructions Executed"]] or 0)
    s = int(r[ci["# Samples"]] or 0)
    th = int(r[ci["Thread Instructions Executed"]] or 0)
    a = agg[(f, ln)]
    a[0] += n; a[1] += s; a[2] += th
    if wf_i is not None:
        w = int(r[wf_i] or 0); a[4] += w; tw += w
        a[5] += int(r[wfi_i] or 0) if wfi_i is not None else 0
    for i in stall_cols:
        v = r[i]
        if v and v != "0":
            a[3][hdr[i]] += int(v)
    ti += n; ts += s
print(f"total warp-instructions {ti}  samples {ts}  shared-memory wavefronts {tw}")
src_cache = {}
def src(f, ln):
    for base in ("opfgym_b200/csrc", "include"):
        p = os.path.join(base, f)
        if os.path.exists(p):
            if p not in src_cache:
                src_cache[p] = open(p).read().splitlines()
            L = src_cache[p]
            return L[ln - 1].strip()[:90] if 0 < ln <= len(L) else ""
    return ""
order = {"samples": 1, "inst": 0, "wavefronts": 4}[sort_key]
for (f, ln), (n, s, th, st, w, wi) in sorted(agg.items(), key=lambda kv: -kv[1][order])[:top]:
    lanes = th / n if n else 0
    tops = ",".join(f"{k[6:]}:{v}" for k, v in st.most_common(3))
    wf = f" {w/tw*100:5.1f}%wf(x{w/wi:.2f})" if tw and wi else (f" {w/tw*100:5.1f}%wf" if tw else "")
    print(f"{s/ts*100:5.1f}%smp {n/ti*100:5.1f}%inst{wf} lanes={lanes:4.1f} {f}:{ln:<4} {src(f, ln)}   [{tops}]")
