#!/bin/bash
# Quick GPU pass: parity tests, one bench line, optional ncu capture of the index kernel (NCU=1).
set -u
mkdir -p gpurun_out
if [ "${SKIP_TESTS:-0}" != "1" ]; then
timeout -k 10 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_gpu.log
fi
timeout -k 10 400 python bench.py "$@" > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -5 gpurun_out/bench.err
cat gpurun_out/bench.json
if [ "${NCU:-0}" = "1" ]; then
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k "regex:stream_kernel|tile_scan|tile_gather" -s 9 -c 3 -f -o gpurun_out/prof_index \
   python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e "$@" > gpurun_out/ncu_full.log 2>&1; echo "ncu full index rc=$?"
fi
