#!/bin/bash
# Round-end GPU visit: bench line, ncu launch list of the bench, full captures of the association and training kernels.
# usage: scripts/gpu_final.sh <tag>
tag=${1:-r3f}
mkdir -p gpurun_out
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
tail -2 gpurun_out/${tag}_bench.err; cut -c1-400 gpurun_out/${tag}_bench.json
GENIE_BENCH_PROFILE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 300 --csv \
  --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --day-seconds 2000 > gpurun_out/${tag}_ncu_bench.log 2>&1
tail -2 gpurun_out/${tag}_ncu_bench.log
cd scripts
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'assoc_init|assoc_layer|src_mean' -c 10 \
  -o ../gpurun_out/${tag}_assoc_prof -f python probe_assoc.py c4s > ../gpurun_out/${tag}_ncu_assoc.log 2>&1
tail -2 ../gpurun_out/${tag}_ncu_assoc.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'kron_spmm' -c 6 \
  -o ../gpurun_out/${tag}_train_prof -f python probe_train.py > ../gpurun_out/${tag}_ncu_train.log 2>&1
tail -2 ../gpurun_out/${tag}_ncu_train.log
ls -la ../gpurun_out/ | tail -8
