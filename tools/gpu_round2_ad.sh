#!/bin/bash
# GPU call AD (1 GPU): validation of the final round-2 code -- full -m gpu suite, smoke, the default bench (driver's command), the launch list.
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/ad_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/ad_pytest.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/ad_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/ad_smoke.log
( time timeout 900 python bench.py ) > gpurun_out/ad_bench.json 2> gpurun_out/ad_bench.err; echo "bench rc=$?" >> gpurun_out/ad_bench.err
CMD="python bench.py --proof-k 20 --log-n 0 --ntt-log-n 0 --batch-k 0 --mst-log-n 0 --no-checker --no-cpu-baseline --steps 1 --warmup 3"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/ad_launches_k20.csv $CMD > gpurun_out/ad_ncu1.log 2>&1
echo done
