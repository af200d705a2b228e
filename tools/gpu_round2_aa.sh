#!/bin/bash
# GPU call AA (2 GPUs): replicated lagrange_to_coeff transforms on the side stream, under the Lagrange-basis commitments.
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q ) > gpurun_out/aa_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/aa_pytest.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
( timeout 900 $T --master-port 29621 bench.py --gpus 2 --steps 5 --warmup 3 --log-n 0 --ntt-log-n 0 --batch-k 0 --mst-log-n 0 --no-cpu-baseline ) > gpurun_out/aa_bench_n2.json 2> gpurun_out/aa_bench_n2.err; echo "rc=$?" >> gpurun_out/aa_bench_n2.err
( timeout 900 python bench.py --steps 5 --warmup 3 --log-n 0 --ntt-log-n 0 --batch-k 0 --mst-log-n 0 --no-cpu-baseline ) > gpurun_out/aa_bench_n1.json 2> gpurun_out/aa_bench_n1.err; echo "rc=$?" >> gpurun_out/aa_bench_n1.err
echo done
