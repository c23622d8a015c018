#!/bin/bash
# targeted GPU tests (pytest -k expression in $1)
set -u
mkdir -p gpurun_out
timeout -k 10 ${TEST_TIMEOUT:-420} python -m pytest tests -m gpu -x -q -k "$1" > gpurun_out/pytest_k.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/pytest_k.log
