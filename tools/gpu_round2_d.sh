#!/bin/bash
# GPU call D of round 2 (2 GPUs): boundary entries, BigUint tree, distributed four-step NTT, sharded proof on 2 GPUs, 2-GPU bench.
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/d_pytest.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 3 --warmup 3 ) > gpurun_out/d_bench_n2.json 2> gpurun_out/d_bench_n2.err; echo "bench rc=$?" >> gpurun_out/d_bench_n2.err
echo done
