#!/bin/bash
# Round 2, wide kernel rework II: parity tests, default bench line, knob variants, instruction counts, e2e probe.
set -u
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log
line() { python - "$1" <<'PY'
import sys, json
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("step %.3f ms idx %.3f kern_frac %.3f build_frac %.3f q %.3f q_frac %.3f rows %d cpu_match %s" % (
        d["ms_per_step"], d["index_ms"], d["roofline"]["frac"], d["roofline_index_build"]["frac"], d["query_ms"],
        d["roofline_query"]["frac"], d["index_rows"], (d.get("cpu_baseline") or {}).get("matches_gpu")))
except Exception as e:
    print("ERR", e)
PY
}
timeout -k 10 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-extras > gpurun_out/w4_bench.json 2> gpurun_out/w4_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/w4_bench.err; line gpurun_out/w4_bench.json
i=0
for env in "MEMO_WIDE_PREFETCH=0" "MEMO_WIDE_STRIP_ROWS=920" "MEMO_WIDE_STRIP_ROWS=230" "MEMO_WIDE_STRIP_ROWS=1840"; do
  i=$((i+1))
  timeout -k 10 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-extras --env $env > gpurun_out/w4_bench_v$i.json 2> gpurun_out/w4_bench_v$i.err; echo "$env rc=$?"
  line gpurun_out/w4_bench_v$i.json
done
timeout -k 10 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum \
  --clock-control none -k regex:"wide_kernel|strip_gather" -c 6 --csv --log-file gpurun_out/w4_ncu_wide.csv \
  python bench.py --rows 10000000 --steps 1 --warmup 1 --no-cpu --no-e2e --no-extras > gpurun_out/w4_ncu.log 2>&1; echo "ncu rc=$?"
grep -E "wide_kernel|strip_gather" gpurun_out/w4_ncu_wide.csv | awk -F'","' '{print $2, $5, $(NF-2), $(NF)}' | sed 's/void unnamed>:://' | cut -c1-160
timeout -k 10 300 python scripts/e2e_probe93.py > gpurun_out/w4_e2e_probe.log 2>&1; echo "probe rc=$?"
cat gpurun_out/w4_e2e_probe.log | tail -12
