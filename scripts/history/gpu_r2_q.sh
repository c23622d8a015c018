#!/bin/bash
# query heavy-tile threshold experiment on a first shard (record start included) of chr1/8 x 94
set -u
mkdir -p gpurun_out
for h in default 256 512 1024 4096 0; do
E=""; if [ "$h" != "default" ]; then E="--env MEMO_QUERY_HEAVY=$h"; fi
timeout -k 10 200 python bench.py --rows 31119553 --cols 93 --no-cpu --no-e2e --no-extras $E > gpurun_out/bq.json 2> gpurun_out/bq.err; echo "heavy=$h rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/bq.json"))
print("  query %.4f ms frac %.3f" % (d["query_ms"], d["roofline_query"]["frac"]))
PY
done
