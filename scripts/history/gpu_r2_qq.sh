#!/bin/bash
set -u
mkdir -p gpurun_out
timeout -k 10 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "query or golden or cli" > gpurun_out/pytest_q.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -3 gpurun_out/pytest_q.log
if [ $rc -ne 0 ]; then exit 1; fi
for cfg in "--rows 100000000 --cols 9" "--rows 10000000 --cols 93" ""; do
timeout -k 10 300 python bench.py --no-cpu --no-e2e --no-extras $cfg > gpurun_out/bi.json 2> gpurun_out/bi.err; echo "[$cfg] rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/bi.json"))
print("  step %.3f idx %.3f query %.4f qfrac %.3f" % (d["ms_per_step"], d["index_ms"], d["query_ms"], d["roofline_query"]["frac"]))
PY
done
