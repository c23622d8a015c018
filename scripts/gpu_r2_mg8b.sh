#!/bin/bash
# 8-GPU pass (short): strong-scaled chr1 x 94 bench with shard parity
set -u
N=${N:-8}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_g$N.json 2> gpurun_out/bench_g$N.err; echo "bench rc=$?"
tail -2 gpurun_out/bench_g$N.err | cut -c1-300
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_g$N.json").read().strip().splitlines()[-1])
print("N=%d: step %.3f ms idx %.3f kern %.3f query %.3f frac %.3f build %.3f q %.3f parity %s e2e %.1f Mbp/s h2d %.1f GB/s/gpu (ceiling %.1f)" % (d["n_gpus"], d["ms_per_step"], d["index_ms"], d["roofline"]["kernel_ms"], d["query_ms"], d["roofline"]["frac"], d["roofline_index_build"]["frac"], d["roofline_query"]["frac"], d["shard_parity_detail"], d["e2e"]["value"]/1e6, d["e2e"]["h2d_gbs_per_gpu"], d["e2e"]["h2d_ceiling_gbs_per_gpu"]))
PY
