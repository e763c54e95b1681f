#!/bin/bash
# Final round-2 evidence after the merge-on-mma.sync and precision="bf16" changes: full GPU suite, bench (all extras), ncu launch list,
# full captures of every hot kernel (same launch indices as scratch/r2_evidence.sh) + the new ones.  Run under gpurun; scratch/make_profiles.py gpurun_out/r02ev2 r02 afterwards.
cd $GRAFT_REPO_ROOT
o=gpurun_out/r02ev2; mkdir -p $o
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > $o/smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > $o/pytest.log 2>&1; echo rc=$? >> $o/pytest.log; tail -3 $o/pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > $o/bench.json 2> $o/bench.err; tail -2 $o/bench.err
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > $o/bench_reference.json 2> $o/bench_reference.err; tail -2 $o/bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $o/launches.csv python scratch/prof_rollout.py 128 1 > $o/ncu_launch.log 2>&1
cap() {  # name regex skip [precision]
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:$2 -s $3 -c 1 -o $o/$1 -f python scratch/prof_rollout.py 128 1 ${4:-bf16x3} > $o/ncu_$1.log 2>&1
  ncu -i $o/$1.ncu-rep --page raw --csv > $o/$1.raw.csv
  ncu -i $o/$1.ncu-rep --page source --csv > $o/$1.source.csv
  rm -f $o/$1.ncu-rep
}
cap merge k_merge 20
cap merge_late k_merge 40
cap derive k_node_derive 0
cap rowqk_bf16 k_tc_gemm 2 bf16
cap rowpv_bf16 k_tc_gemm 3 bf16
cap rowqk k_tc_gemm2s 2
cap rowpv k_tc_gemm2w 1
# the other 11 captures (score / alpha / column block / ffn / rowqkv / softmax) are those of scratch/r2_evidence.sh: kernels unchanged since
ls -la $o | tail -50
