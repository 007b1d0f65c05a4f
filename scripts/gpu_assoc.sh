#!/bin/bash
# GPU visit for the association branch: its parity tests first, then the whole GPU suite, then a bench line.
# usage: scripts/gpu_assoc.sh <tag>
tag=${1:-r3}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "forward_fixed_matches or forward_equals or association_matches" > gpurun_out/${tag}_assoc_pytest.log 2>&1
echo "assoc pytest rc=$?" >> gpurun_out/${tag}_assoc_pytest.log
tail -40 gpurun_out/${tag}_assoc_pytest.log
timeout 600 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_parity.py::test_forward_fixed_matches_reference --deselect tests/test_gpu_parity.py::test_forward_equals_forward_fixed_and_refuses_training --deselect tests/test_gpu_parity.py::test_association_matches_oracle_seeded > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
tail -3 gpurun_out/${tag}_bench.err; cat gpurun_out/${tag}_bench.json
