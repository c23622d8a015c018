#!/bin/bash
# why is configs[1] slower inside the full bench line than alone?
set -u
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import sys, json
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
for x in d.get("extra_configs") or []:
    print("   ", x["config"][:40], "idx %.3f kern %.3f (%.3f)" % (x["index_ms"], x["roofline"]["kernel_ms"], x["roofline"]["frac"]))
PY
}
timeout -k 10 600 python bench.py --steps 5 --no-e2e --no-cpu > gpurun_out/x1.json 2> gpurun_out/x1.err; echo "no-e2e no-cpu rc=$?"; show gpurun_out/x1.json
timeout -k 10 600 python bench.py --steps 5 --no-cpu > gpurun_out/x2.json 2> gpurun_out/x2.err; echo "no-cpu rc=$?"; show gpurun_out/x2.json
timeout -k 10 600 python bench.py --steps 5 --no-e2e > gpurun_out/x3.json 2> gpurun_out/x3.err; echo "no-e2e rc=$?"; show gpurun_out/x3.json
