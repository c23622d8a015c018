#!/bin/bash
# index kernels: parity tests, then 10 Mbp x 93 and full-size timing
set -u
mkdir -p gpurun_out
timeout -k 10 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "index or golden or stream or smoke" > gpurun_out/pytest_idx.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -4 gpurun_out/pytest_idx.log
if [ $rc -ne 0 ]; then exit 1; fi
for cfg in "--rows 10000000 --cols 93" "--membership --rows 5000000 --cols 93" ""; do
timeout -k 10 300 python bench.py --no-cpu --no-e2e --no-extras $cfg > gpurun_out/bi.json 2> gpurun_out/bi.err; echo "[$cfg] rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/bi.json"))
print("  step %.3f idx %.3f kern %.3f frac %.3f build %.3f q %.3f" % (d["ms_per_step"], d["index_ms"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["roofline_index_build"]["frac"], d["roofline_query"]["frac"]))
PY
done
if [ "${NCU:-0}" = "1" ]; then
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k "regex:wide_kernel" -s 4 -c 1 -f -o gpurun_out/prof_wide \
   python bench.py --rows 10000000 --cols 93 --steps 2 --warmup 3 --no-cpu --no-e2e --no-extras > gpurun_out/ncu_wide.log 2>&1; echo "ncu rc=$?"
fi
