#!/bin/bash
# prefetch modes: time (bench) and DRAM traffic of the strip kernel (ncu) at chr1 x 94
set -u
mkdir -p gpurun_out
for m in 0 1 2 3; do
  timeout -k 10 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-extras --env MEMO_WIDE_PREFETCH=$m > gpurun_out/w10_b$m.json 2> gpurun_out/w10_b$m.err; echo "mode $m rc=$?"
  python - <<PY
import json
d = json.loads(open("gpurun_out/w10_b$m.json").read().strip().splitlines()[-1])
print("  step %.3f idx %.3f kern %.3f frac %.3f" % (d["ms_per_step"], d["index_ms"], d["roofline"]["kernel_ms"], d["roofline"]["frac"]))
PY
  timeout -k 10 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum --clock-control none -k regex:wide_kernel -s 2 -c 1 --csv --log-file gpurun_out/w10_n$m.csv \
     python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-extras --env MEMO_WIDE_PREFETCH=$m > gpurun_out/w10_n$m.log 2>&1; echo "  ncu rc=$?"
  grep wide_kernel gpurun_out/w10_n$m.csv | awk -F'","' '{print "   ", $(NF-2), $NF}'
done
