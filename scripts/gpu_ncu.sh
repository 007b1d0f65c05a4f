#!/bin/bash
# ncu --set full of the library kernels matching a regex, inside bench.py's timed loop.  usage: gpu_ncu.sh <tag> <regex> [count]
tag=$1; re=$2; cnt=${3:-4}
mkdir -p gpurun_out
GENIE_BENCH_PROFILE=1 timeout 1500 ncu --profile-from-start off --set full --clock-control none --import-source on \
  -k regex:"$re" -c $cnt -o gpurun_out/${tag}_prof -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --day-seconds 2000 > gpurun_out/${tag}_ncu_full.log 2>&1
tail -3 gpurun_out/${tag}_ncu_full.log
