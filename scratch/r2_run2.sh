#!/bin/bash
# round-2 run #2: full GPU test suite (new breadth tests), smoke, new bench line
mkdir -p gpurun_out/r2
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest2.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2/smoke2.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2/smoke2.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r2/bench2.json 2> gpurun_out/r2/bench2.err; echo "bench rc=$?" >> gpurun_out/r2/bench2.err
tail -5 gpurun_out/r2/pytest2.log
