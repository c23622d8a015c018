#!/bin/bash
# Round-2 evidence, final code state: default bench line, launch list, DRAM traffic of one chr1 x 94 step
set -u
mkdir -p gpurun_out
timeout -k 10 600 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
K="regex:narrow_kernel|wide_kernel|tile_scan|strip_gather|query_planes"
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-extras > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout -k 10 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k "$K" -s 16 -c 4 --csv --log-file gpurun_out/r02_traffic_chr1_x94.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-extras > gpurun_out/ncu_traffic.log 2>&1; echo "ncu traffic rc=$?"
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k "$K" -s 16 -c 4 -f -o gpurun_out/r02_step_c93 \
     python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-extras --cols 93 --rows 10000000 > gpurun_out/ncu_full_c93.log 2>&1; echo "ncu full c93 rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench.json").read().strip().splitlines()[-1])
print("step %.3f ms idx %.3f kern %.3f kern_frac %.3f build_frac %.3f q %.3f q_frac %.3f" % (
    d["ms_per_step"], d["index_ms"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["roofline_index_build"]["frac"], d["query_ms"], d["roofline_query"]["frac"]))
e = d.get("e2e") or {}; t = d.get("e2e_text") or {}
print("e2e %.1f Mbp/s; e2e_text %.2f Mbp/s; clocks %s" % (e.get("value", 0) / 1e6, t.get("value", 0) / 1e6, d["clocks"]))
PY
