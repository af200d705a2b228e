#!/bin/bash
# GPU call B of round 2: the rewritten NTT (register radix-8 tiles, cp.async, fused scalings): parity, variant sweep, bench, ncu.
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/b_pytest.log
( time timeout 600 python tools/ntt_sweep.py ) > gpurun_out/b_ntt_sweep.jsonl 2> gpurun_out/b_ntt_sweep.err
( time timeout 900 python bench.py ) > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err; echo "bench rc=$?" >> gpurun_out/b_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ntt_pass_kernel -s 6 -c 4 -o gpurun_out/b_ntt_full -f \
  python bench.py --proof-k 20 --log-n 0 --ntt-log-n 22 --batch-k 0 --mst-log-n 0 --no-checker --steps 1 --warmup 3 > gpurun_out/b_ncu_ntt.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:msm_reduce_first_kernel -s 12 -c 2 -o gpurun_out/b_msm_l1_proof_full -f \
  python bench.py --proof-k 20 --log-n 0 --ntt-log-n 0 --batch-k 0 --mst-log-n 0 --no-checker --steps 1 --warmup 3 > gpurun_out/b_ncu_msm.log 2>&1
echo done
