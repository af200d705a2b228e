#!/bin/bash
# GPU call Q (1 GPU): dedicated Montgomery squaring in the curve arithmetic, the Poseidon S-boxes and the generated evaluate_h -- parity, then timings.
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/q_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/q_pytest.log
( timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline ) > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err; echo "rc=$?" >> gpurun_out/q_bench.err
echo done
