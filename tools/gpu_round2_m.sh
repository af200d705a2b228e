#!/bin/bash
# GPU call M (1 GPU): tree bucket reduction, run-aggregated sort, two-level batch inversion, direct instance cosets -- parity, then A/B timings.
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/test_gpu_msm.py tests/test_gpu_prover.py tests/test_gpu_baseline_k.py -m gpu -x -q ) > gpurun_out/m_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/m_pytest.log
B="python bench.py --steps 5 --warmup 3 --proof-k 17,20 --log-n 22 --ntt-log-n 0 --batch-k 0 --mst-log-n 0 --no-cpu-baseline"
run() { name=$1; shift; ( env "$@" timeout 600 $B ) > gpurun_out/m_bench_$name.json 2> gpurun_out/m_bench_$name.err; echo "rc=$?" >> gpurun_out/m_bench_$name.err; }
run default SB_X=1
run no_tree SB_MSM_NO_BUCKET_TREE=1
run no_binv2 SB_NO_BINV2=1
run no_inst SB_NO_INST_DIRECT=1
run cta64k SB_MSM_CTA_SCAN_MAX=65536
run cta256k SB_MSM_CTA_SCAN_MAX=262144
echo done
