#!/bin/bash
# Round 2: host path without stalls -- parity tests, the driver's default bench line, e2e probe.
set -u
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log
t0=$(date +%s)
timeout -k 10 900 python bench.py > gpurun_out/w5_bench.json 2> gpurun_out/w5_bench.err; echo "bench rc=$? ($(( $(date +%s) - t0 )) s)"
tail -3 gpurun_out/w5_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/w5_bench.json").read().strip().splitlines()[-1])
print("step %.3f ms idx %.3f kern_frac %.3f build_frac %.3f q %.3f q_frac %.3f cpu_match %s" % (
    d["ms_per_step"], d["index_ms"], d["roofline"]["frac"], d["roofline_index_build"]["frac"], d["query_ms"],
    d["roofline_query"]["frac"], (d.get("cpu_baseline") or {}).get("matches_gpu")))
e = d.get("e2e") or {}
print("e2e %.1f Mbp/s %.1f ms/step h2d %.1f GB/s (ceiling %.1f)" % (e.get("value", 0) / 1e6, e.get("ms_per_step", 0), e.get("h2d_gbs_per_gpu", 0), e.get("h2d_ceiling_gbs_per_gpu", 0)))
print("e2e_text", json.dumps(d.get("e2e_text"))[:300])
for k, v in (d.get("extra_configs") or {}).items():
    print(k, json.dumps(v)[:400])
print("gpu_launches", d.get("gpu_launches"), "clocks", d.get("clocks"))
PY
timeout -k 10 300 python scripts/e2e_probe93.py > gpurun_out/w5_e2e_probe.log 2>&1; echo "probe rc=$?"
tail -8 gpurun_out/w5_e2e_probe.log
