#!/bin/bash
# bench lines + the captures that involve the wide kernel (after gpu_r2_final.sh)
set -u
mkdir -p gpurun_out
timeout -k 10 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout -k 10 600 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
timeout -k 10 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
K="regex:narrow_kernel|wide_kernel|wide2_kernel|wide2_finish|tile_scan|strip_gather|query_planes"
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-extras > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout -k 10 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k "$K" -s 16 -c 4 --csv --log-file gpurun_out/r02_traffic_chr1_x94.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-extras > gpurun_out/ncu_traffic.log 2>&1; echo "ncu traffic rc=$?"
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k "$K" -s 16 -c 4 -f -o gpurun_out/r02_step_c93 \
     python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-extras --cols 93 --rows 10000000 > gpurun_out/ncu_full_c93.log 2>&1; echo "ncu full c93 rc=$?"
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k "$K" -s 16 -c 4 -f -o gpurun_out/r02_step_memb \
     python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-extras --membership --cols 93 --rows 5000000 > gpurun_out/ncu_full_memb.log 2>&1; echo "ncu full memb rc=$?"
