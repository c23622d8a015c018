#!/usr/bin/env python3
"""DRAM traffic of one index build from an `ncu --set full` report of one bench step:
sum of dram__bytes_read.sum + dram__bytes_write.sum over the build's kernels.
usage: ncu_traffic.py report.ncu-rep key [out.json]   (key e.g. index_build_c9)"""
import csv, io, json, os, subprocess, sys
rep, key = sys.argv[1], sys.argv[2]
out = sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.path.dirname(__file__), "..", "profiles", "traffic.json")
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
h, u = rows[0], rows[1]
col = {n: i for i, n in enumerate(h)}
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
total, seen = 0.0, []
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    if not any(k in name for k in ("narrow_kernel", "wide_kernel", "tile_scan", "tile_gather", "strip_gather")):
        continue
    b = 0.0
    for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        b += float(r[col[m]]) * scale[u[col[m]]]
    seen.append((name.split("(")[0][-40:], b, r[col["gpu__time_duration.sum"]] + " " + u[col["gpu__time_duration.sum"]]))
    total += b
for s in seen:
    print("%-42s %14.0f B  %s" % s)
data = {}
if os.path.exists(out):
    data = json.load(open(out))
data[key] = total
data[key + "_kernels"] = [{"kernel": a, "dram_bytes": b, "duration_under_ncu": c} for a, b, c in seen]
json.dump(data, open(out, "w"), indent=1)
print(key, total)
