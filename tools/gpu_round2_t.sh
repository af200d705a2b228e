#!/bin/bash
# GPU call T (1 GPU): level-1 chunk / upper-level chunk sweep after the tail changes (k = 17 and 20 proofs + the 2^22 MSM record).
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
B="python bench.py --steps 5 --warmup 3 --proof-k 17,20 --log-n 22 --ntt-log-n 0 --batch-k 0 --mst-log-n 0 --no-cpu-baseline --no-checker"
run() { name=$1; shift; ( env "$@" timeout 600 $B ) > gpurun_out/t_bench_$name.json 2> gpurun_out/t_bench_$name.err; echo "rc=$?" >> gpurun_out/t_bench_$name.err; }
run default SB_X=1
run l1_32 SB_MSM_L1=32
run l1_128 SB_MSM_L1=128
run lk4 SB_MSM_LK=4
run lk16 SB_MSM_LK=16
run cta16k SB_MSM_CTA_SCAN_MAX=16384
run cta1m SB_MSM_CTA_SCAN_MAX=1048576
echo done
