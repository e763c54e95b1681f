#!/bin/bash
# A/B of library builds on the same box: usage r2_ab.sh <tag> <lib1> <lib2> ...   (libs relative to repo root; "cur" = neuralnj_b200/libnnj.so)
mkdir -p gpurun_out/r2
cd $GRAFT_REPO_ROOT
tag=$1; shift
timeout 900 python -m pytest tests/test_gpu_tc_paths.py tests/test_gpu_parity.py -m gpu -x -q -k "not 200x4096" > gpurun_out/r2/ab_${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/ab_${tag}_pytest.log
tail -3 gpurun_out/r2/ab_${tag}_pytest.log
for rep in 1 2; do
for lib in "$@"; do
  if [ "$lib" = "cur" ]; then p=neuralnj_b200/libnnj.so; else p=$lib; fi
  echo "== $lib" >> gpurun_out/r2/ab_${tag}.jsonl
  NNJ_LIB_PATH=$GRAFT_REPO_ROOT/$p timeout 300 python scratch/r2_explore.py 128 50 1024 bf16x3 3 >> gpurun_out/r2/ab_${tag}.jsonl 2>> gpurun_out/r2/ab_${tag}.err
done
done
