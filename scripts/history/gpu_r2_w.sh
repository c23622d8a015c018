#!/bin/bash
# wide-kernel occupancy experiment: warps per CTA sweep on the current library
set -u
mkdir -p gpurun_out
for w in 6 7 8; do
timeout -k 10 200 python bench.py --rows 10000000 --cols 93 --no-cpu --no-e2e --no-extras --warps $w > gpurun_out/bw.json 2> gpurun_out/bw.err; echo "warps $w rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/bw.json"))
print("10M warps $w: idx %.3f kern %.3f frac %.3f build %.3f" % (d["index_ms"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["roofline_index_build"]["frac"]))
PY
done
for w in 6 8; do
timeout -k 10 300 python bench.py --no-cpu --no-e2e --no-extras --warps $w > gpurun_out/bw.json 2> gpurun_out/bw.err; echo "full warps $w rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/bw.json"))
print("full warps $w: step %.3f idx %.3f kern %.3f frac %.3f build %.3f" % (d["ms_per_step"], d["index_ms"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["roofline_index_build"]["frac"]))
PY
done
