#!/bin/bash
# Short GPU visit while developing: parity tests (stop at first failure) + timing probe.  usage: gpu_quick.sh <tag> [probe args]
tag=${1:-q}; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -30 gpurun_out/${tag}_pytest.log
timeout 600 python scripts/probe_perf.py "$@" > gpurun_out/${tag}_probe.log 2>&1
cat gpurun_out/${tag}_probe.log | tail -40
