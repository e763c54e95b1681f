#!/bin/bash
# Round-2 evidence: bench (with cpu baseline), search bench, chunk sweep, ncu launch list and full captures of the top kernels.
cd $GRAFT_REPO_ROOT
tag=r02
o=gpurun_out/r02ev; mkdir -p $o
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > $o/smi.txt 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 > $o/bench.json 2> $o/bench.err; tail -2 $o/bench.err
timeout 600 python bench.py --workload config4 --steps 3 --warmup 3 > $o/bench_config4.json 2> $o/bench_config4.err; tail -2 $o/bench_config4.err
timeout 600 python scratch/bench_search.py > $o/search.json 2> $o/search.err; cat $o/search.json
for ch in 2 4 8 16 32 64 128; do
  NNJ_CHUNK_MAX=$ch timeout 300 python scratch/r2_explore.py 128 50 1024 bf16x3 2 >> $o/chunk_sweep.jsonl 2>> $o/chunk_sweep.err
done
NNJ_LIB_PATH=$GRAFT_REPO_ROOT/scratch/libnnj_trace.so timeout 300 python scratch/col_trace.py 32 50 1024 > $o/col_trace.txt 2>&1
nvcc -gencode arch=compute_100a,code=sm_100a -O3 scratch/hmma_bench.cu -o scratch/hmma_bench && ./scratch/hmma_bench > $o/hmma_bench.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $o/launches.csv python scratch/prof_rollout.py 128 1 > $o/ncu_launch.log 2>&1
cap() {  # name regex skip
  ncu --set full --import-source on --clock-control none -k regex:$2 -s $3 -c 1 -o $o/$1 -f python scratch/prof_rollout.py 128 1 > $o/ncu_$1.log 2>&1
  ncu -i $o/$1.ncu-rep --page raw --csv > $o/$1.raw.csv
  ncu -i $o/$1.ncu-rep --page source --csv > $o/$1.source.csv
  rm -f $o/$1.ncu-rep
}
cap score_incr k_score_inc 15
cap score_late k_score_inc 30
cap score_small k_score_small 4
cap score_step0 k_score_tc 0
cap alpha_incr k_alpha_v3 20
cap alpha_late k_alpha_v3 36
cap alpha_small k_alpha_small 4
cap colblock k_enc_colblock 2
cap ffn k_enc_ffn 2
cap rowqkv k_enc_rowqkv 2
cap softmax k_softmax_rows 2
cap rowqk k_tc_gemm 2
cap rowpv k_tc_gemm 3
cap merge k_merge 20
ls -la $o | tail -40
