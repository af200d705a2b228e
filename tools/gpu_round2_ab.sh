#!/bin/bash
# GPU call AB (2 GPUs): grand products (ratios, inversion, running products) sharded by row range.
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q ) > gpurun_out/ab_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/ab_pytest.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
B="bench.py --gpus 2 --steps 5 --warmup 3 --log-n 0 --ntt-log-n 0 --batch-k 0 --mst-log-n 0 --no-cpu-baseline"
( timeout 900 $T --master-port 29631 $B ) > gpurun_out/ab_bench_n2.json 2> gpurun_out/ab_bench_n2.err; echo "rc=$?" >> gpurun_out/ab_bench_n2.err
( SB_NO_GRAND_SHARD=1 timeout 900 $T --master-port 29632 $B ) > gpurun_out/ab_bench_n2_off.json 2> gpurun_out/ab_bench_n2_off.err; echo "rc=$?" >> gpurun_out/ab_bench_n2_off.err
echo done
