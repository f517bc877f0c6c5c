#!/usr/bin/env python
"""DRAM traffic of the dominant power-flow kernel from an ncu capture -> profiles/pf_traffic.json.

usage: ncu_traffic.py <report.ncu-rep> <config-name> <n_env> [kernel-regex]
bench.py reads the file and reports `roofline.traffic` only while the hash of csrc/ recorded here
still matches the sources it runs (a stale capture reads as null, never as a number)."""
import csv, hashlib, io, json, os, re, subprocess, sys

rep, config, n_env = sys.argv[1], sys.argv[2], int(sys.argv[3])
pattern = sys.argv[4] if len(sys.argv) > 4 else r"k_pf_(tree|multi|lanes)"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
best = None
for r in rows[2:]:
    if not re.search(pattern, r[col["Kernel Name"]]):
        continue
    rd = float(r[col["dram__bytes_read.sum"]]) * scale[units[col["dram__bytes_read.sum"]]]
    wr = float(r[col["dram__bytes_write.sum"]]) * scale[units[col["dram__bytes_write.sum"]]]
    ms = float(r[col["gpu__time_duration.sum"]]) * {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}[units[col["gpu__time_duration.sum"]]]
    best = dict(kernel=r[col["Kernel Name"]], dram_read=rd, dram_write=wr, duration_ms=ms)
if best is None:
    sys.exit("no kernel matching " + pattern)
h = hashlib.sha256()
for name in ("opfg_core.h", "opfg_api.cu", "symbolic.cpp", "symbolic.hpp"):
    h.update(open(os.path.join(root, "opfgym_b200", "csrc", name), "rb").read())
path = os.path.join(root, "profiles", "pf_traffic.json")
data = json.load(open(path)) if os.path.exists(path) else {}
data[config] = dict(kernel=best["kernel"], n_env=n_env, dram_bytes_per_launch=best["dram_read"] + best["dram_write"],
                    dram_read=best["dram_read"], dram_write=best["dram_write"], duration_ms_under_ncu=best["duration_ms"],
                    csrc_sha256=h.hexdigest(), source=f"ncu --set full capture {os.path.basename(rep)} "
                    f"(dram read {best['dram_read']/1e6:.2f} + write {best['dram_write']/1e6:.2f} MB per launch of {n_env} envs)")
json.dump(data, open(path, "w"), indent=1)
print(json.dumps(data[config], indent=1))
