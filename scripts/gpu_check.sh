#!/bin/bash
# One GPU-box pass: parity tests, bench line, ncu launch list, ncu full capture of the index kernel.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
cat gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:index_kernel -s 3 -c 1 -f -o gpurun_out/prof_index \
   python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_full.log 2>&1; echo "ncu full index rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:query_conservation -s 3 -c 1 -f -o gpurun_out/prof_query \
   python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_full_q.log 2>&1; echo "ncu full query rc=$?"
ls -la gpurun_out
