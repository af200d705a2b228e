#!/bin/bash
# GPU call N (1 GPU): full -m gpu suite on the new MSM tail / batch inversion / instance path, then the launch list of one k = 20 proof.
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/n_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/n_pytest.log
CMD="python bench.py --proof-k 20 --log-n 0 --ntt-log-n 0 --batch-k 0 --mst-log-n 0 --no-checker --no-cpu-baseline --steps 1 --warmup 3"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/n_launches_k20.csv $CMD > gpurun_out/n_ncu1.log 2>&1
echo done
