#!/bin/bash
# Round 2: tests + default bench line + launch list (gather as one resident wave, device BED formatter)
set -u
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log
timeout -k 10 900 python bench.py > gpurun_out/w8_bench.json 2> gpurun_out/w8_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/w8_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/w8_bench.json").read().strip().splitlines()[-1])
print("step %.3f ms idx %.3f kern_frac %.3f build_frac %.3f q %.3f q_frac %.3f cpu_match %s" % (
    d["ms_per_step"], d["index_ms"], d["roofline"]["frac"], d["roofline_index_build"]["frac"], d["query_ms"],
    d["roofline_query"]["frac"], (d.get("cpu_baseline") or {}).get("matches_gpu")))
e = d.get("e2e") or {}
print("e2e %.1f Mbp/s %.1f ms/step h2d %.1f GB/s (ceiling %.1f)" % (e.get("value", 0) / 1e6, e.get("ms_per_step", 0), e.get("h2d_gbs_per_gpu", 0), e.get("h2d_ceiling_gbs_per_gpu", 0)))
t = d.get("e2e_text") or {}
print("e2e_text %.2f Mbp/s %.3f s text %.2f GB/s" % (t.get("value", 0) / 1e6, t.get("seconds", 0), t.get("text_gbs", 0)))
for x in d.get("extra_configs") or []:
    print(x["config"][:44], "step %.3f idx %.3f kern %.3f (%.3f) q %.4f" % (x["ms_per_step"], x["index_ms"], x["roofline"]["kernel_ms"], x["roofline"]["frac"], x["query_ms"]), x.get("clocks", {}).get("reasons"))
print("gpu_launches", d.get("gpu_launches"), "clocks", d.get("clocks"))
PY
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/w8_launches.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-extras > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
grep -E "wide_kernel|strip_gather|tile_scan|prep_kernel|query_planes" gpurun_out/w8_launches.csv | tail -10 | awk -F'","' '{print $5, $NF}' | cut -c1-120
