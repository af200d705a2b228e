#!/bin/bash
# GPU call A of round 2: full -m gpu suite, smoke, default bench, reference arm at k=17 (bounded), launch list of the headline proof.
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
nproc > gpurun_out/nproc.txt; free -g >> gpurun_out/nproc.txt; nvidia-smi --query-gpu=name,memory.total --format=csv >> gpurun_out/nproc.txt
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/a_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/a_smoke.log
( time timeout 900 python bench.py ) > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; echo "bench rc=$?" >> gpurun_out/a_bench.err
( time timeout 600 python bench.py --impl reference --ref-k 17 ) > gpurun_out/a_ref17.json 2> gpurun_out/a_ref17.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/a_launches_k20.csv \
  python bench.py --proof-k 20 --log-n 0 --ntt-log-n 0 --batch-k 0 --mst-log-n 0 --no-checker --steps 1 --warmup 3 > gpurun_out/a_ncu.log 2>&1
echo done
