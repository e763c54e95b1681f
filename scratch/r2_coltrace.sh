#!/bin/bash
mkdir -p gpurun_out/r2
cd $GRAFT_REPO_ROOT
tag=${1:-col}
NNJ_LIB_PATH=$GRAFT_REPO_ROOT/scratch/libnnj_trace.so timeout 300 python scratch/col_trace.py 32 50 1024 > gpurun_out/r2/${tag}_trace.txt 2>&1
cat gpurun_out/r2/${tag}_trace.txt
timeout 1500 python -m pytest tests -m gpu -x -q -k "not 200x4096" > gpurun_out/r2/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/${tag}_pytest.log
tail -5 gpurun_out/r2/${tag}_pytest.log
timeout 300 python scratch/r2_explore.py 128 50 1024 bf16x3 3 >> gpurun_out/r2/${tag}.jsonl 2>> gpurun_out/r2/${tag}.err
timeout 300 python scratch/r2_explore.py 128 20 256 bf16x3 3 >> gpurun_out/r2/${tag}.jsonl 2>> gpurun_out/r2/${tag}.err
timeout 300 python scratch/r2_explore.py 32 100 1024 bf16x3 2 >> gpurun_out/r2/${tag}.jsonl 2>> gpurun_out/r2/${tag}.err
cat gpurun_out/r2/${tag}.jsonl; tail -5 gpurun_out/r2/${tag}.err
