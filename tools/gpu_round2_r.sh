#!/bin/bash
# GPU call R (2 GPUs): distributed four-step NTT for the replicated lagrange_to_coeff transforms from k = 17 up (default: k >= 22) -- A/B.
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
B="bench.py --gpus 2 --steps 3 --warmup 3 --proof-k 17,20 --log-n 0 --ntt-log-n 0 --batch-k 0 --mst-log-n 0 --no-cpu-baseline"
( SB_DIST_NTT_MIN_K=17 timeout 900 $T --master-port 29581 $B ) > gpurun_out/r_bench_n2_dist17.json 2> gpurun_out/r_bench_n2_dist17.err; echo "rc=$?" >> gpurun_out/r_bench_n2_dist17.err
( timeout 900 $T --master-port 29582 $B ) > gpurun_out/r_bench_n2_default.json 2> gpurun_out/r_bench_n2_default.err; echo "rc=$?" >> gpurun_out/r_bench_n2_default.err
echo done
