#!/bin/bash
# Round-2 closing pass on the committed code: smoke, GPU tests, both bench arms
set -u
mkdir -p gpurun_out
timeout -k 10 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout -k 10 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02_pytest_gpu.log
timeout -k 10 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout -k 10 600 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench.json").read().strip().splitlines()[-1])
print("step %.3f ms idx %.3f kern %.3f kern_frac %.3f build_frac %.3f q %.3f q_frac %.3f" % (
    d["ms_per_step"], d["index_ms"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["roofline_index_build"]["frac"], d["query_ms"], d["roofline_query"]["frac"]))
e = d.get("e2e") or {}; t = d.get("e2e_text") or {}
print("e2e %.1f Mbp/s; e2e_text %.2f Mbp/s; clocks %s" % (e.get("value", 0) / 1e6, t.get("value", 0) / 1e6, d["clocks"]))
for x in d.get("extra_configs") or []:
    print("   ", x["config"][:40], "step %.3f idx %.3f (%.3f) kern %.3f (%.3f) q %.4f (%.3f)" % (x["ms_per_step"], x["index_ms"], x["roofline_index_build"]["frac"], x["roofline"]["kernel_ms"], x["roofline"]["frac"], x["query_ms"], x["roofline_query"]["frac"]))
PY
