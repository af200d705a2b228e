#!/bin/bash
# GPU call P (2 GPUs): sharded proofs after the evaluation-domain SHPLONK (row-sharded), sharded evaluations, tree bucket reduction.
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q ) > gpurun_out/p_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/p_pytest.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 3 --warmup 3 --log-n 0 --ntt-log-n 0 --batch-k 0 --mst-log-n 0 --no-cpu-baseline ) > gpurun_out/p_bench_n2.json 2> gpurun_out/p_bench_n2.err; echo "rc=$?" >> gpurun_out/p_bench_n2.err
( SB_NO_SHPLONK_SHARD=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29572 bench.py --gpus 2 --steps 3 --warmup 3 --log-n 0 --ntt-log-n 0 --batch-k 0 --mst-log-n 0 --no-cpu-baseline ) > gpurun_out/p_bench_n2_noshard.json 2> gpurun_out/p_bench_n2_noshard.err; echo "rc=$?" >> gpurun_out/p_bench_n2_noshard.err
echo done
