#!/bin/bash
set -u
mkdir -p gpurun_out
timeout -k 5 120 compute-sanitizer --tool memcheck --print-limit 5 python scripts/w2_repro.py > gpurun_out/sanitizer.log 2>&1; echo "sanitizer rc=$?"
grep -v "^=========     Host Frame\|^=========         in\|^=========                in" gpurun_out/sanitizer.log | head -60
