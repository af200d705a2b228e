#!/bin/bash
# GPU call G (8 GPUs): the bench exactly as the driver launches it at N = 8 and N = 4, and the 2-GPU sharded tests.
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 3 --warmup 3 ) > gpurun_out/g_bench_n8.json 2> gpurun_out/g_bench_n8.err; echo "bench8 rc=$?" >> gpurun_out/g_bench_n8.err
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 4 --steps 3 --warmup 3 ) > gpurun_out/g_bench_n4.json 2> gpurun_out/g_bench_n4.err; echo "bench4 rc=$?" >> gpurun_out/g_bench_n4.err
( time timeout 600 python bench.py --impl reference --gpus 1 ) > gpurun_out/g_ref_k20.json 2> gpurun_out/g_ref_k20.err
nproc > gpurun_out/g_nproc.txt
echo done
