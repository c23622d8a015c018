#!/bin/bash
# text path: read threads 1 vs 4 vs 8 (e2e_text leg of a small bench)
set -u
mkdir -p gpurun_out
for th in 1 4 8 1 8; do
timeout -k 10 300 python bench.py --rows 20000000 --steps 3 --warmup 3 --no-cpu --no-extras --e2e-steps 1 --e2e-rows 2000000 --e2e-text-rows 2000000 --env MEMO_TEXT_READ_THREADS=$th > gpurun_out/w9_$th.json 2> gpurun_out/w9_$th.err; echo "threads=$th rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/w9_$th.json").read().strip().splitlines()[-1])
t = d.get("e2e_text") or {}
print("e2e_text %.2f Mbp/s %.3f s text %.2f GB/s" % (t.get("value", 0) / 1e6, t.get("seconds", 0), t.get("text_gbs", 0)))
PY
done
nproc; free -g | head -2
