#!/bin/bash
# quick gate + 10 Mbp x 93 timing of both strip kernels (+ optional ncu) + full-size bench
set -u
mkdir -p gpurun_out
timeout -k 5 90 python scripts/w2_repro.py > gpurun_out/repro.log 2>&1; rc=$?; echo "repro rc=$rc"; tail -3 gpurun_out/repro.log
if [ $rc -ne 0 ] || grep -q "equal False" gpurun_out/repro.log; then echo "STOP: repro failed"; exit 1; fi
for v in 0 2; do
timeout -k 10 200 python bench.py --rows 10000000 --cols 93 --no-cpu --no-e2e --no-extras --variant $v > gpurun_out/bench_c93_v$v.json 2> gpurun_out/bench_c93_v$v.err; echo "variant $v rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_c93_v$v.json"))
print("variant $v: step %.3f ms idx %.3f kern %.3f query %.3f frac %.3f build %.3f" % (d["ms_per_step"], d["index_ms"], d["roofline"]["kernel_ms"], d["query_ms"], d["roofline"]["frac"], d["roofline_index_build"]["frac"]))
PY
done
if [ "${NCU:-0}" = "1" ]; then
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k "regex:wide2_kernel" -s 4 -c 1 -f -o gpurun_out/prof_w2 \
   python bench.py --rows 10000000 --cols 93 --steps 2 --warmup 3 --no-cpu --no-e2e --no-extras > gpurun_out/ncu_w2.log 2>&1; echo "ncu rc=$?"
fi
if [ "${FULL:-0}" = "1" ]; then
timeout -k 10 400 python bench.py --no-extras --no-cpu --no-e2e "$@" > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "full rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_full.json"))
print("full: step %.3f ms idx %.3f kern %.3f query %.3f frac %.3f build %.3f qfrac %.3f" % (d["ms_per_step"], d["index_ms"], d["roofline"]["kernel_ms"], d["query_ms"], d["roofline"]["frac"], d["roofline_index_build"]["frac"], d["roofline_query"]["frac"]))
PY
fi
