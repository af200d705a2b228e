#!/bin/bash
# GPU call X (1 GPU): validation of the final round-2 code -- full -m gpu suite, smoke, the default bench (driver's command), the reference arm at
# k = 17 (bounded), the launch list of one k = 20 proof and --set full captures of the kernels that changed (bucket tree, sort, level 1, evaluate_h).
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
nproc > gpurun_out/x_box.txt; nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv >> gpurun_out/x_box.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/x_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/x_pytest.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/x_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/x_smoke.log
( time timeout 900 python bench.py ) > gpurun_out/x_bench.json 2> gpurun_out/x_bench.err; echo "bench rc=$?" >> gpurun_out/x_bench.err
CMD="python bench.py --proof-k 20 --log-n 0 --ntt-log-n 0 --batch-k 0 --mst-log-n 0 --no-checker --no-cpu-baseline --steps 1 --warmup 3"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/x_launches_k20.csv $CMD > gpurun_out/x_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:msm_bucket_tree -s 20 -c 2 -o gpurun_out/x_tree -f $CMD > gpurun_out/x_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:msm_reduce_first -s 21 -c 1 -o gpurun_out/x_msm1 -f $CMD > gpurun_out/x_ncu3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:msm_sort -s 42 -c 2 -o gpurun_out/x_sort -f $CMD > gpurun_out/x_ncu4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sb_h_jit -s 12 -c 1 -o gpurun_out/x_hjit -f $CMD > gpurun_out/x_ncu5.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ntt_pass -s 120 -c 2 -o gpurun_out/x_ntt -f $CMD > gpurun_out/x_ncu6.log 2>&1
echo done
