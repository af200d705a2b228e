#!/bin/bash
# GPU call J (2 GPUs): the library's own communicator (shared-memory mailbox + CUDA IPC) vs the NCCL callbacks.
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q ) > gpurun_out/j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j_pytest.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 3 --warmup 3 --log-n 0 --ntt-log-n 0 --batch-k 0 --mst-log-n 0 ) > gpurun_out/j_bench_n2_shm.json 2> gpurun_out/j_bench_n2_shm.err; echo "rc=$?" >> gpurun_out/j_bench_n2_shm.err
( SB_BENCH_COMM=nccl timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus 2 --steps 3 --warmup 3 --log-n 0 --ntt-log-n 0 --batch-k 0 --mst-log-n 0 ) > gpurun_out/j_bench_n2_nccl.json 2> gpurun_out/j_bench_n2_nccl.err; echo "rc=$?" >> gpurun_out/j_bench_n2_nccl.err
echo done
