#!/bin/bash
# round-2 exploration run #1: baseline bench, chunk-size sweep (L2 residency), small-batch latency, large shapes
mkdir -p gpurun_out/r2
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2/smi.txt 2>&1
python bench.py --steps 3 --warmup 3 > gpurun_out/r2/bench_base.json 2> gpurun_out/r2/bench_base.err
for ch in 4 8 16 32 64; do
  NNJ_CHUNK_MAX=$ch timeout 300 python scratch/r2_explore.py 128 50 1024 bf16x3 2 >> gpurun_out/r2/chunk_sweep.jsonl 2>> gpurun_out/r2/chunk_sweep.err
done
timeout 300 python scratch/r2_explore.py 128 50 1024 bf16x3 2 >> gpurun_out/r2/chunk_sweep.jsonl 2>> gpurun_out/r2/chunk_sweep.err
for b in 1 2 4 8 16; do
  timeout 300 python scratch/r2_explore.py $b 50 1024 bf16x3 5 >> gpurun_out/r2/small_batch.jsonl 2>> gpurun_out/r2/small_batch.err
done
timeout 600 python scratch/r2_explore.py 2 200 4096 bf16x3 1 >> gpurun_out/r2/large.jsonl 2>> gpurun_out/r2/large.err
timeout 600 python scratch/r2_explore.py 32 100 1024 bf16x3 1 >> gpurun_out/r2/large.jsonl 2>> gpurun_out/r2/large.err
timeout 600 python scratch/r2_explore.py 64 50 1024 fp32 1 >> gpurun_out/r2/large.jsonl 2>> gpurun_out/r2/large.err
echo done
