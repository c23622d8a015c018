#!/bin/bash
# Round 2: small shapes after the wide-kernel rework (strip length vs tail; narrow kernel)
set -u
mkdir -p gpurun_out
line() { python - "$1" <<'PY'
import sys, json
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("step %.3f ms idx %.3f kern %.3f kern_frac %.3f q %.4f rows %d clocks %s" % (
        d["ms_per_step"], d["index_ms"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["query_ms"], d["index_rows"], d["clocks"]["reasons"]))
except Exception as e:
    print("ERR", e)
PY
}
i=0
run() { i=$((i+1)); timeout -k 10 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --no-extras "$@" > gpurun_out/w6_$i.json 2> gpurun_out/w6_$i.err; echo "$* rc=$?"; line gpurun_out/w6_$i.json; }
run --rows 100000000 --cols 9
run --rows 100000000 --cols 9
for sr in 115 230 460; do run --rows 10000000 --cols 93 --env MEMO_WIDE_STRIP_ROWS=$sr; done
for sr in 115 230 460; do run --rows 5000000 --cols 93 --membership --env MEMO_WIDE_STRIP_ROWS=$sr; done
run --rows 10000000 --cols 93 --env MEMO_WIDE_STRIP_ROWS=115 --env MEMO_WIDE_PREFETCH=0
