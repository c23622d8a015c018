#!/bin/bash
# Round-2 evidence, last pass: default bench line (traffic.json refreshed), ncu --set full of the strip
# kernel and the gather at chr1 x 94
set -u
mkdir -p gpurun_out
timeout -k 10 600 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k "regex:wide_kernel|strip_gather" -s 8 -c 2 -f -o gpurun_out/r02_build_chr1_x94 \
     python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-extras > gpurun_out/ncu_full_chr1.log 2>&1; echo "ncu full chr1 rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench.json").read().strip().splitlines()[-1])
print("step %.3f ms idx %.3f kern %.3f kern_frac %.3f build_frac %.3f q %.3f q_frac %.3f traffic %.4g / %.4g" % (
    d["ms_per_step"], d["index_ms"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["roofline_index_build"]["frac"], d["query_ms"], d["roofline_query"]["frac"], d["roofline"]["traffic"], d["roofline_index_build"]["traffic"]))
e = d.get("e2e") or {}; t = d.get("e2e_text") or {}
print("e2e %.1f Mbp/s; e2e_text %.2f Mbp/s; clocks %s" % (e.get("value", 0) / 1e6, t.get("value", 0) / 1e6, d["clocks"]))
PY
ls -la gpurun_out/r02_build_chr1_x94.ncu-rep
