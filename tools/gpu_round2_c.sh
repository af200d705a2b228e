#!/bin/bash
# GPU call C of round 2: 5-coset quotient pipeline + batched NTT launches: parity, bench.
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c_pytest.log
( time timeout 900 python bench.py ) > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err; echo "bench rc=$?" >> gpurun_out/c_bench.err
( time timeout 300 python tools/ntt_sweep.py 20 22 ) > gpurun_out/c_ntt_sweep.jsonl 2> gpurun_out/c_ntt_sweep.err
echo done
