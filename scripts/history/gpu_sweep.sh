#!/bin/bash
# Tuning sweep of the index build (bench.py knobs); prints index_ms per configuration.
set -u
mkdir -p gpurun_out
: > gpurun_out/sweep.txt
run() {
  out=$(timeout -k 10 300 python bench.py --no-cpu --no-e2e --steps 10 --warmup 3 "$@" 2>&1 | tail -1)
  echo "$* => $(echo "$out" | python -c 'import sys,json
try:
    d=json.loads(sys.stdin.read()); print("index_ms=%.3f query_ms=%.3f frac=%.3f replays=%s" % (d["index_ms"], d["query_ms"], d["roofline"]["frac"], d["replayed_strips"]))
except Exception as e: print("ERR", e)')" | tee -a gpurun_out/sweep.txt
}
while read -r line; do
  [ -z "$line" ] && continue
  run $line
done < "${1:-scripts/sweep_configs.txt}"
