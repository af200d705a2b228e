#!/bin/bash
# GPU call O (1 GPU): SHPLONK in the evaluation domain + binary-Euclid inversion -- parity (full suite), then A/B timings.
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/o_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/o_pytest.log
B="python bench.py --steps 5 --warmup 3 --proof-k 17,20,23 --log-n 0 --ntt-log-n 0 --batch-k 0 --mst-log-n 0 --no-cpu-baseline"
run() { name=$1; shift; ( env "$@" timeout 900 $B ) > gpurun_out/o_bench_$name.json 2> gpurun_out/o_bench_$name.err; echo "rc=$?" >> gpurun_out/o_bench_$name.err; }
run default SB_X=1
run no_lagr SB_NO_SHPLONK_LAGRANGE=1
echo done
