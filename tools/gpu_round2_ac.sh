#!/bin/bash
# GPU call AC (8 GPUs): the driver's scaling command on the final code, plus the k = 20 proof without the row-sharded grand products (A/B).
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
( time timeout 900 $T --master-port 29641 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline ) > gpurun_out/ac_bench_n8.json 2> gpurun_out/ac_bench_n8.err; echo "rc=$?" >> gpurun_out/ac_bench_n8.err
( SB_NO_GRAND_SHARD=1 timeout 600 $T --master-port 29642 bench.py --gpus 8 --steps 5 --warmup 3 --proof-k 17,20 --log-n 0 --ntt-log-n 0 --batch-k 0 --mst-log-n 0 --no-cpu-baseline --no-checker ) > gpurun_out/ac_bench_n8_off.json 2> gpurun_out/ac_bench_n8_off.err; echo "rc=$?" >> gpurun_out/ac_bench_n8_off.err
echo done
