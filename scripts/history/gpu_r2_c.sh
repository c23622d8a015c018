#!/bin/bash
set -u
mkdir -p gpurun_out
timeout -k 10 200 python bench.py --rows 10000000 --cols 93 --no-cpu --no-e2e --no-extras > gpurun_out/bench_c93_v0.json 2> gpurun_out/bench_c93_v0.err; echo "rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_c93_v0.json"))
print("10M: step %.3f ms idx %.3f kern %.3f query %.3f frac %.3f build %.3f parked %d of %d" % (d["ms_per_step"], d["index_ms"], d["roofline"]["kernel_ms"], d["query_ms"], d["roofline"]["frac"], d["roofline_index_build"]["frac"], d["index_rows_parked_rank0"], d["index_rows"]))
PY
timeout -k 10 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_c93.csv python bench.py --rows 10000000 --cols 93 --steps 2 --warmup 3 --no-cpu --no-e2e --no-extras > /dev/null 2>&1
grep -v "^==" gpurun_out/launches_c93.csv | tail -12 | cut -d, -f5,12-
