#!/bin/bash
# Iteration pass: GPU parity tests, then one bench line per argument set in $1 (a file, one set per line).
set -u
mkdir -p gpurun_out
if [ "${SKIP_TESTS:-0}" != "1" ]; then
timeout -k 10 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_gpu.log
fi
: > gpurun_out/iter.txt
i=0
while read -r line; do
  [ -z "$line" ] && continue
  i=$((i+1))
  out=$(timeout -k 10 400 python bench.py $line 2> gpurun_out/iter_$i.err | tail -1)
  echo "$out" > gpurun_out/iter_$i.json
  echo "$line => $(echo "$out" | python -c 'import sys,json
try:
    d=json.loads(sys.stdin.read()); print("value=%.3e index_ms=%.3f query_ms=%.3f frac=%.3f rows=%d e2e=%s" % (d["value"], d["index_ms"], d["query_ms"], d["roofline"]["frac"], d["index_rows"], (d.get("e2e") or {}).get("value")))
except Exception as e: print("ERR", e)')" | tee -a gpurun_out/iter.txt
  tail -2 gpurun_out/iter_$i.err
done < "${1:-scripts/iter_configs.txt}"
