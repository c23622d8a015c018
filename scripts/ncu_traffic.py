#!/usr/bin/env python3
"""DRAM traffic per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the kernels of one
bench step, from an `ncu --set full` report (.ncu-rep) or a `--metrics ... --csv` log, into
profiles/traffic.json under the keys bench.py looks up:
    <tag>_stream_kernel   the streaming kernel of the index build
    <tag>_index_build     every kernel of the build
    <tag>_query           the query kernel
usage: ncu_traffic.py report.ncu-rep|log.csv tag [out.json]   (tag e.g. c93_cons_248956422)"""
import csv, io, json, os, subprocess, sys
rep, tag = sys.argv[1], sys.argv[2]
out = sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.path.dirname(__file__), "..", "profiles", "traffic.json")
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6, "usecond": 1e3, "msecond": 1e6, "nsecond": 1}
per = []          # [(kernel, bytes, duration ns)] in launch order
if rep.endswith(".ncu-rep"):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    h, u = rows[0], rows[1]
    col = {n: i for i, n in enumerate(h)}
    for r in rows[2:]:
        b = sum(float(r[col[m]]) * scale[u[col[m]]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        d = float(r[col["gpu__time_duration.sum"]]) * scale[u[col["gpu__time_duration.sum"]]]
        per.append((r[col["Kernel Name"]], b, d))
else:
    rows = [r for r in csv.reader(l for l in open(rep) if not l.startswith("==")) if len(r) > 5]
    h = rows[0]
    col = {n: i for i, n in enumerate(h)}
    acc = {}
    for r in rows[1:]:
        key = r[col["ID"]]
        name, metric, unit, val = r[col["Kernel Name"]], r[col["Metric Name"]], r[col["Metric Unit"]], float(r[col["Metric Value"]])
        e = acc.setdefault(key, [name, 0.0, 0.0])
        if metric.startswith("dram__bytes"):
            e[1] += val * scale[unit]
        elif metric == "gpu__time_duration.sum":
            e[2] = val * scale[unit]
    per = [tuple(v) for v in acc.values()]
data = json.load(open(out)) if os.path.exists(out) else {}
build = [p for p in per if any(k in p[0] for k in ("narrow_kernel", "wide_kernel", "wide2_", "tile_scan", "strip_gather"))]
stream = [p for p in build if any(k in p[0] for k in ("narrow_kernel", "wide_kernel", "wide2_kernel"))]
query = [p for p in per if "query_" in p[0]]
if stream:
    data[f"{tag}_stream_kernel"] = stream[0][1]
if build:
    # one build = the kernels up to the next streaming kernel
    n = len(build) // max(len(stream), 1)
    data[f"{tag}_index_build"] = sum(p[1] for p in build[:n])
    data[f"{tag}_index_build_kernels"] = [{"kernel": p[0].split("(")[0][-40:], "dram_bytes": p[1], "duration_under_ncu_ns": p[2]} for p in build[:n]]
if query:
    data[f"{tag}_query"] = query[0][1]
    data[f"{tag}_query_duration_under_ncu_ns"] = query[0][2]
data["_source"] = "ncu captures of round 2 (scripts/gpu_r2_final2.sh, gpu_r2_final3.sh; --clock-control none), per launch; looked up by bench.py, not measured in its run"
json.dump(data, open(out, "w"), indent=1)
for k in (f"{tag}_stream_kernel", f"{tag}_index_build", f"{tag}_query"):
    print(k, data.get(k))
