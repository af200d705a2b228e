#!/bin/bash
# GPU call Y4 (4 GPUs): the driver's scaling command on the final round-2 code.
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29614 bench.py --gpus 4 --steps 5 --warmup 3 --no-cpu-baseline ) > gpurun_out/y_bench_n4.json 2> gpurun_out/y_bench_n4.err; echo "rc=$?" >> gpurun_out/y_bench_n4.err
echo done
