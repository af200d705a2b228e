#!/bin/bash
# GPU call K (1 GPU): 4-elements-per-thread NTT tiles (SB_NTT_EB=2: 512 threads per 2^11 tile, 32 warps / SM) against the radix-8 default.
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for eb in 3 2; do
  ( SB_NTT_EB=$eb timeout 600 python tools/ntt_sweep.py 16 18 20 22 24 ) > gpurun_out/k_ntt_sweep_eb$eb.txt 2> gpurun_out/k_ntt_sweep_eb$eb.err; echo "rc=$?" >> gpurun_out/k_ntt_sweep_eb$eb.err
  ( SB_NTT_EB=$eb timeout 600 python bench.py --steps 5 --warmup 3 --proof-k 20 --log-n 0 --ntt-log-n 22 --batch-k 0 --mst-log-n 0 --no-cpu-baseline ) > gpurun_out/k_bench_eb$eb.json 2> gpurun_out/k_bench_eb$eb.err; echo "rc=$?" >> gpurun_out/k_bench_eb$eb.err
done
echo done
