#!/bin/bash
mkdir -p gpurun_out/r2
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2/final_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/final_pytest.log
tail -5 gpurun_out/r2/final_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
