#!/bin/bash
# GPU call H: blocking-sync worker contexts, evaluate_h operand staging, level-1 chunk sweep at k = 20.
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/h_pytest.log
( time timeout 1200 python bench.py ) > gpurun_out/h_bench.json 2> gpurun_out/h_bench.err; echo "bench rc=$?" >> gpurun_out/h_bench.err
for l1 in 128 256; do
  ( SB_MSM_L1=$l1 timeout 300 python bench.py --proof-k 20 --log-n 0 --ntt-log-n 0 --batch-k 0 --mst-log-n 0 --no-checker --steps 3 ) > gpurun_out/h_l1_$l1.json 2> gpurun_out/h_l1_$l1.err
done
( SB_MSM_LK=16 timeout 300 python bench.py --proof-k 20 --log-n 0 --ntt-log-n 0 --batch-k 0 --mst-log-n 0 --no-checker --steps 3 ) > gpurun_out/h_lk_16.json 2> gpurun_out/h_lk_16.err
echo done
