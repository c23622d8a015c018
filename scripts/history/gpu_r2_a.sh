#!/bin/bash
# Round-2 pass A: quick repro, smoke, parity tests, default bench line (chr1 x 94), reference arm.
# Every stage is gated on the one before it (a hung kernel must not burn the GPU budget).
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/smi.txt 2>&1
timeout -k 5 90 python scripts/w2_repro.py > gpurun_out/repro.log 2>&1; rc=$?; echo "repro rc=$rc"; tail -3 gpurun_out/repro.log
if [ $rc -ne 0 ] || grep -q "equal False" gpurun_out/repro.log; then echo "STOP: repro failed"; exit 1; fi
timeout -k 10 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; rc=$?; echo "smoke rc=$rc"; tail -2 gpurun_out/smoke.log
if [ $rc -ne 0 ]; then echo "STOP: smoke failed"; exit 1; fi
if [ "${SKIP_TESTS:-0}" != "1" ]; then
timeout -k 10 ${TEST_TIMEOUT:-420} python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest rc=$rc"
tail -8 gpurun_out/pytest_gpu.log
if [ $rc -ne 0 ]; then echo "STOP: tests failed"; exit 1; fi
fi
( time timeout -k 10 400 python bench.py "$@" > gpurun_out/bench.json 2> gpurun_out/bench.err ) 2> gpurun_out/bench.time; echo "bench rc=$?"
tail -5 gpurun_out/bench.err; cat gpurun_out/bench.time
cat gpurun_out/bench.json
if [ "${SKIP_REF:-0}" != "1" ]; then
timeout -k 10 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
cat gpurun_out/bench_ref.json
fi
