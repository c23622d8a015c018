#!/bin/bash
# 8-GPU pass: strong-scaled chr1 x 94 bench with shard parity, then the whole-genome sweep
set -u
N=${N:-8}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1; free -g > gpurun_out/mem8.txt; nproc >> gpurun_out/mem8.txt
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_g$N.json 2> gpurun_out/bench_g$N.err; echo "bench rc=$?"
tail -2 gpurun_out/bench_g$N.err | cut -c1-300
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_g$N.json").read().strip().splitlines()[-1])
print("N=%d: step %.3f ms idx %.3f kern %.3f query %.3f frac %.3f build %.3f q %.3f parity %s e2e %.1f Mbp/s h2d %.1f GB/s/gpu" % (d["n_gpus"], d["ms_per_step"], d["index_ms"], d["roofline"]["kernel_ms"], d["query_ms"], d["roofline"]["frac"], d["roofline_index_build"]["frac"], d["roofline_query"]["frac"], d["shard_parity_detail"], d["e2e"]["value"]/1e6, d["e2e"]["h2d_gbs_per_gpu"]))
PY
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus $N --wg > gpurun_out/bench_wg_g$N.json 2> gpurun_out/bench_wg_g$N.err; echo "wg rc=$?"
tail -2 gpurun_out/bench_wg_g$N.err | cut -c1-300
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_wg_g$N.json").read().strip().splitlines()[-1])
print("WG N=%d: %.1f Gbp/s total %.1f ms idx %.1f ms (frac %.3f, kernel %.3f) sweep %.1f ms (frac %.3f) rows %d" % (d["n_gpus"], d["value"]/1e9, d["ms_total"], d["index_ms"], d["roofline_index_build"]["frac"], d["roofline"]["frac"], d["query_ms"], d["roofline_query"]["frac"], d["index_rows"]))
PY
