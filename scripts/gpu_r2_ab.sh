#!/bin/bash
# A/B in one box: libmemo_b200_A.so (experiment build) against libmemo_b200.so; $AARGS = bench args of the A runs
set -u
mkdir -p gpurun_out
L=memo_b200/csrc
cp $L/libmemo_b200.so /tmp/B.so; cp $L/libmemo_b200_A.so /tmp/A.so
one() { timeout -k 10 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-extras "${@:2}" > gpurun_out/ab_$1.json 2> gpurun_out/ab_$1.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/ab_$1.json").read().strip().splitlines()[-1])
    print("  $1: step %.3f idx %.3f kern %.3f frac %.3f" % (d["ms_per_step"], d["index_ms"], d["roofline"]["kernel_ms"], d["roofline"]["frac"]))
except Exception as e:
    print("  $1: ERR", e); print(open("gpurun_out/ab_$1.err").read()[-600:])
PY
}
for r in 1 2; do
  cp /tmp/A.so $L/libmemo_b200.so; one A$r ${AARGS:-}
  cp /tmp/B.so $L/libmemo_b200.so; one B$r
done
cp /tmp/A.so $L/libmemo_b200.so; one A10m --rows 10000000 ${AARGS:-}
cp /tmp/B.so $L/libmemo_b200.so; one B10m --rows 10000000
