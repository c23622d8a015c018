#!/bin/bash
# ncu --set full capture of kernels matching $1 (regex) in a short bench run with the remaining args; report -> gpurun_out/$NAME.ncu-rep
set -u
mkdir -p gpurun_out
name=${NAME:-prof}
rx=$1; shift
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k "regex:$rx" -s ${SKIP:-3} -c ${COUNT:-1} -f -o gpurun_out/$name \
   python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e "$@" > gpurun_out/$name.log 2>&1; echo "ncu $name rc=$?"
tail -3 gpurun_out/$name.log
