#!/bin/bash
mkdir -p gpurun_out/r2
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -x -q -k "config4 or w100 or t100 or 200x4096" > gpurun_out/r2/big_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/big_pytest.log
tail -30 gpurun_out/r2/big_pytest.log
rm -f gpurun_out/r2/big.jsonl
timeout 600 python scratch/r2_explore.py 2 200 4096 bf16x3 1 >> gpurun_out/r2/big.jsonl 2>> gpurun_out/r2/big.err
timeout 600 python scratch/r2_explore.py 4 200 4096 bf16x3 1 >> gpurun_out/r2/big.jsonl 2>> gpurun_out/r2/big.err
timeout 600 python scratch/r2_explore.py 32 100 1024 bf16x3 1 >> gpurun_out/r2/big.jsonl 2>> gpurun_out/r2/big.err
cat gpurun_out/r2/big.jsonl; tail -5 gpurun_out/r2/big.err
