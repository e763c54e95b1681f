#!/bin/bash
mkdir -p gpurun_out/r2
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests/test_gpu_llh.py -m gpu -x -q > gpurun_out/r2/llh_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/llh_pytest.log
tail -40 gpurun_out/r2/llh_pytest.log
