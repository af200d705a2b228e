#!/bin/bash
# GPU call L (1 GPU): ncu evidence of the final code -- launch list of the headline proof, --set full of evaluate_h (NVRTC) and MSM level 1 inside the proof.
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
CMD="python bench.py --proof-k 20 --log-n 0 --ntt-log-n 0 --batch-k 0 --mst-log-n 0 --no-checker --no-cpu-baseline --steps 1 --warmup 3"
( timeout 600 $CMD ) > gpurun_out/l_bench.json 2> gpurun_out/l_bench.err; echo "rc=$?" >> gpurun_out/l_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/l_launches_k20.csv $CMD > gpurun_out/l_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sb_h_jit -s 12 -c 1 -o gpurun_out/l_hjit -f $CMD > gpurun_out/l_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:msm_reduce_first -s 20 -c 2 -o gpurun_out/l_msm1 -f $CMD > gpurun_out/l_ncu3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:msm_reduce_kernel -s 60 -c 4 -o gpurun_out/l_msm2 -f $CMD > gpurun_out/l_ncu4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:msm_sort -s 40 -c 2 -o gpurun_out/l_sort -f $CMD > gpurun_out/l_ncu5.log 2>&1
ls -la gpurun_out
echo done
