#!/bin/bash
# warps per CTA (chunk rows follow the shared memory per warp)
set -u
mkdir -p gpurun_out
i=0
run() { i=$((i+1)); timeout -k 10 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-extras "$@" > gpurun_out/w12_$i.json 2> gpurun_out/w12_$i.err
python - "$*" <<PY
import sys, json
try:
    d = json.loads(open("gpurun_out/w12_$i.json").read().strip().splitlines()[-1])
    print("%-50s step %.3f idx %.3f kern %.3f frac %.3f" % (sys.argv[1], d["ms_per_step"], d["index_ms"], d["roofline"]["kernel_ms"], d["roofline"]["frac"]))
except Exception as e:
    print(sys.argv[1], "ERR", e, open("gpurun_out/w12_$i.err").read()[-300:])
PY
}
run
run --warps 5
run --warps 5 --rows-per-tile 26
run --warps 4 --rows-per-tile 23
run
