#!/bin/bash
# Development visit: parity tests, then the front-end timing probe at C4 for each value of an environment knob.
# usage: scripts/gpu_sweep.sh <tag> <ENV_NAME> <v1> <v2> ...
tag=$1; var=$2; shift 2
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -15 gpurun_out/${tag}_pytest.log
for v in "$@"; do
  echo "== $var=$v" | tee -a gpurun_out/${tag}_sweep.log
  env $var=$v timeout 300 python scripts/probe_perf.py c4 2>&1 | grep -E "front end|kernel " | tee -a gpurun_out/${tag}_sweep.log
done
