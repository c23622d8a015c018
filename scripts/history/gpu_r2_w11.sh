#!/bin/bash
# strip length at the shard sizes of N = 4 / 8 (62 M / 31 M rows per GPU)
set -u
mkdir -p gpurun_out
i=0
run() { i=$((i+1)); timeout -k 10 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --no-extras "$@" > gpurun_out/w11_$i.json 2> gpurun_out/w11_$i.err
python - "$*" <<PY
import sys, json
d = json.loads(open("gpurun_out/w11_$i.json").read().strip().splitlines()[-1])
print("%-60s step %.3f idx %.3f kern %.3f frac %.3f" % (sys.argv[1], d["ms_per_step"], d["index_ms"], d["roofline"]["kernel_ms"], d["roofline"]["frac"]))
PY
}
for rows in 31119553 62239106; do
  for sr in 161 230 299 345 460; do run --rows $rows --env MEMO_WIDE_STRIP_ROWS=$sr; done
done
