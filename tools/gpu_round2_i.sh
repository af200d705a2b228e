#!/bin/bash
# GPU call I: NVRTC-specialised evaluate_h: parity, timing vs the interpreter.
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/i_pytest.log
( time timeout 600 python bench.py --proof-k 17,20,23 --log-n 0 --ntt-log-n 0 --batch-k 0 --mst-log-n 0 --no-checker --steps 3 ) > gpurun_out/i_jit.json 2> gpurun_out/i_jit.err
( SB_NO_JIT=1 timeout 600 python bench.py --proof-k 17,20,23 --log-n 0 --ntt-log-n 0 --batch-k 0 --mst-log-n 0 --no-checker --steps 3 ) > gpurun_out/i_nojit.json 2> gpurun_out/i_nojit.err
echo done
