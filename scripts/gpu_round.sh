#!/bin/bash
# One GPU-box visit: parity tests, timing probe, bench line, ncu launch list + full capture of the top kernels.
# usage: scripts/gpu_round.sh <tag> [skip-ncu|skip-tests|ncu-only]
# The .ncu-rep files stay on the box (gpurun brings back at most 64 MiB): their raw pages come back as CSV.
tag=${1:-r1}
mkdir -p gpurun_out
export_rep() { ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null; rm -f gpurun_out/$1.ncu-rep; }
if [ "$2" != "skip-tests" ] && [ "$2" != "ncu-only" ]; then
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
fi
if [ "$2" != "ncu-only" ]; then
timeout 600 python scripts/probe_perf.py c2 c4 > gpurun_out/${tag}_probe.log 2>&1
cat gpurun_out/${tag}_probe.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
tail -3 gpurun_out/${tag}_bench.err; cat gpurun_out/${tag}_bench.json
fi
if [ "$2" != "skip-ncu" ]; then
GENIE_BENCH_PROFILE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 300 --csv \
  --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity-check --day-seconds 2000 > gpurun_out/${tag}_ncu_bench.log 2>&1
# --set full + the tensor-pipe counters BASELINE.json's north star names
GENIE_BENCH_PROFILE=1 timeout 1200 ncu --profile-from-start off --set full \
  --metrics sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_uniform.sum \
  --clock-control none \
  -k regex:'input_gather|da_init|src_mean|da_layer1_s|da_layer2_s|window_' -c 7 -o gpurun_out/${tag}_prof -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity-check --day-seconds 2000 > gpurun_out/${tag}_ncu_full.log 2>&1
tail -3 gpurun_out/${tag}_ncu_full.log
export_rep ${tag}_prof
# the same capture of the bf16-storage mode's kernels (bench.py's second mode)
GENIE_BENCH_PROFILE=bf16 timeout 1200 ncu --profile-from-start off --set full \
  --metrics sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_uniform.sum \
  --clock-control none \
  -k regex:'da_init|src_mean|da_layer1_s|da_layer2_s' -c 5 -o gpurun_out/${tag}_prof_bf16 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity-check --day-seconds 2000 > gpurun_out/${tag}_ncu_full_bf16.log 2>&1
tail -3 gpurun_out/${tag}_ncu_full_bf16.log
export_rep ${tag}_prof_bf16
# association branch (scripts/probe_assoc.py at 1000 x 5000): one forward_fixed = the front end's two station passes, then
# assoc_init and the ASSOC instances of the two station-pass kernels
GENIE_PROBE_ASSOC_ONLY=1 timeout 900 ncu --set full --metrics sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum \
  --clock-control none -k regex:'assoc_init|da_layer1_s|da_layer2_s' -c 5 \
  -o gpurun_out/${tag}_prof_assoc -f python scripts/probe_assoc.py c4s > gpurun_out/${tag}_ncu_assoc.log 2>&1
tail -3 gpurun_out/${tag}_ncu_assoc.log
export_rep ${tag}_prof_assoc
timeout 300 python scripts/probe_assoc.py c2 c4 > gpurun_out/${tag}_assoc_probe.log 2>&1
tail -22 gpurun_out/${tag}_assoc_probe.log
fi
