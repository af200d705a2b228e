#!/bin/bash
# GPU call W (1 GPU): 2^10-element NTT tiles up to 2^20 -- parity of the transforms, then the proofs.
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_prover.py tests/test_gpu_baseline_k.py -m gpu -x -q ) > gpurun_out/w_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/w_pytest.log
( timeout 600 python bench.py --steps 5 --warmup 3 --proof-k 17,20 --log-n 0 --ntt-log-n 20 --batch-k 0 --mst-log-n 0 --no-cpu-baseline ) > gpurun_out/w_bench.json 2> gpurun_out/w_bench.err; echo "rc=$?" >> gpurun_out/w_bench.err
( SB_NTT_TILE=11 timeout 600 python bench.py --steps 5 --warmup 3 --proof-k 17,20 --log-n 0 --ntt-log-n 20 --batch-k 0 --mst-log-n 0 --no-cpu-baseline ) > gpurun_out/w_bench_tile11.json 2> gpurun_out/w_bench_tile11.err; echo "rc=$?" >> gpurun_out/w_bench_tile11.err
echo done
