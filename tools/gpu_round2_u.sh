#!/bin/bash
# GPU call U (2 GPUs): commitments sharded by bucket residue (default) vs by signed-digit window.
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q ) > gpurun_out/u_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/u_pytest.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
B="bench.py --gpus 2 --steps 3 --warmup 3 --log-n 0 --ntt-log-n 0 --batch-k 0 --mst-log-n 0 --no-cpu-baseline"
( timeout 900 $T --master-port 29601 $B ) > gpurun_out/u_bench_n2_residue.json 2> gpurun_out/u_bench_n2_residue.err; echo "rc=$?" >> gpurun_out/u_bench_n2_residue.err
( SB_SHARD_MSM_BY_WINDOW=1 timeout 900 $T --master-port 29602 $B ) > gpurun_out/u_bench_n2_window.json 2> gpurun_out/u_bench_n2_window.err; echo "rc=$?" >> gpurun_out/u_bench_n2_window.err
echo done
