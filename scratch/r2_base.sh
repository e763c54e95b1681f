#!/bin/bash
# session baseline: full GPU suite + default bench line
mkdir -p gpurun_out/r2
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2/base_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/base_pytest.log
tail -5 gpurun_out/r2/base_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2/base_bench.json 2> gpurun_out/r2/base_bench.err
tail -c 3000 gpurun_out/r2/base_bench.json
