#!/bin/bash
# One GPU-box pass: parity tests, bench line (both arms), ncu launch list, ncu full capture of one step's kernels.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout -k 10 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout -k 10 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -3 gpurun_out/bench.err
cat gpurun_out/bench.json
timeout -k 10 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
cat gpurun_out/bench_ref.json
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
# one whole step (4 kernels) after the sizing run and 3 warm-up steps
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k "regex:narrow_kernel|wide_kernel|tile_scan|tile_gather|strip_gather|query_planes|query_conservation|query_membership" -s 15 -c 4 -f -o gpurun_out/prof_step \
   python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k "regex:narrow_kernel|wide_kernel|tile_scan|tile_gather|strip_gather|query_planes|query_conservation|query_membership" -s 15 -c 4 -f -o gpurun_out/prof_step_c93 \
   python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --cols 93 --rows 10000000 > gpurun_out/ncu_full_c93.log 2>&1; echo "ncu full c93 rc=$?"
timeout -k 10 300 python bench.py --no-cpu --no-e2e --cols 93 --rows 10000000 > gpurun_out/bench_c93.json 2> gpurun_out/bench_c93.err; cat gpurun_out/bench_c93.json
ls -la gpurun_out
timeout -k 10 300 python bench.py --no-cpu --no-e2e --membership --cols 93 --rows 5000000 > gpurun_out/bench_memb.json 2> gpurun_out/bench_memb.err; cat gpurun_out/bench_memb.json
