#!/bin/bash
# column-block A/B: parity of the encoder paths, then per-class times at 128 x 50 x 1024
mkdir -p gpurun_out/r2
cd $GRAFT_REPO_ROOT
tag=${1:-col}
timeout 900 python -m pytest tests/test_gpu_tc_paths.py tests/test_gpu_parity.py -m gpu -x -q -k "not 200x4096" > gpurun_out/r2/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/${tag}_pytest.log
tail -15 gpurun_out/r2/${tag}_pytest.log
for rep in 1 2; do
  timeout 300 python scratch/r2_explore.py 128 50 1024 bf16x3 3 >> gpurun_out/r2/${tag}.jsonl 2>> gpurun_out/r2/${tag}.err
done
timeout 300 python scratch/r2_explore.py 128 20 256 bf16x3 3 >> gpurun_out/r2/${tag}.jsonl 2>> gpurun_out/r2/${tag}.err
cat gpurun_out/r2/${tag}.jsonl; tail -5 gpurun_out/r2/${tag}.err
