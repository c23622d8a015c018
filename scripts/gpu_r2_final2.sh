#!/bin/bash
# Round-2 evidence pass (second half of the round: reworked strip kernel, 8-byte scratch rows, resident
# gather, parameter-fed builds, stall-free streaming host path, device BED formatter) on one B200:
# smoke, tests, bench (both arms), ncu launch list + DRAM traffic of the default workload, ncu --set
# full of one step of the smaller shapes, sanitizer runs over the index / formatter tests.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit,memory.total --format=csv > gpurun_out/smi.txt 2>&1
timeout -k 10 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout -k 10 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest_gpu.log
timeout -k 10 600 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
timeout -k 10 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
K="regex:narrow_kernel|wide_kernel|tile_scan|strip_gather|query_planes"
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-extras > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout -k 10 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k "$K" -s 16 -c 4 --csv --log-file gpurun_out/r02_traffic_chr1_x94.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-extras > gpurun_out/ncu_traffic.log 2>&1; echo "ncu traffic rc=$?"
for cfg in "c93 --cols 93 --rows 10000000" "c9 --cols 9 --rows 100000000" "memb --membership --cols 93 --rows 5000000"; do
  set -- $cfg; tag=$1; shift
  timeout -k 10 600 ncu --set full --clock-control none --import-source on -k "$K" -s 16 -c 4 -f -o gpurun_out/r02_step_$tag \
     python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-extras "$@" > gpurun_out/ncu_full_$tag.log 2>&1; echo "ncu full $tag rc=$?"
done
timeout -k 10 420 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "valid_ms_all_geometries or bed_formatter or multi_record or narrow_kernel_shapes" > gpurun_out/r02_index_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/r02_index_memcheck.log
timeout -k 10 420 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "valid_ms_all_geometries or bed_formatter" > gpurun_out/r02_index_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/r02_index_racecheck.log
ls -la gpurun_out/*.ncu-rep
