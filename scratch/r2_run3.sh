#!/bin/bash
mkdir -p gpurun_out/r2
cd $GRAFT_REPO_ROOT
timeout 2000 python -m pytest tests -m gpu -q > gpurun_out/r2/pytest3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest3.log
tail -15 gpurun_out/r2/pytest3.log
