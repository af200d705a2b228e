#!/bin/bash
# GPU call F: full -m gpu suite (1 GPU) after the configs[3]/[4] work, and the c = 17 tables at k = 20.
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f_pytest.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/f_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/f_smoke.log
echo done
