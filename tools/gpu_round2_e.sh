#!/bin/bash
# GPU call E of round 2: configs[3] (<23,8,8>) + configs[4] (distinct users) + witness generator + generic verifier; window-size sweep for the k=20 proof.
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/e_pytest.log
( time timeout 1200 python bench.py ) > gpurun_out/e_bench.json 2> gpurun_out/e_bench.err; echo "bench rc=$?" >> gpurun_out/e_bench.err
for c in 17 18 19; do
  ( SB_TAB_C=$c timeout 300 python bench.py --proof-k 20 --log-n 0 --ntt-log-n 0 --batch-k 0 --mst-log-n 0 --no-checker --steps 3 ) > gpurun_out/e_tabc_$c.json 2> gpurun_out/e_tabc_$c.err
done
echo done
