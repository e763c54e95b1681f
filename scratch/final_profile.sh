#!/bin/bash
# Round-end evidence: tests, bench (with cpu baseline), ncu launch list and full captures of the top kernels.  Run under gpurun.
set -x
tag=${1:-r01}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; tail -2 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -2 gpurun_out/${tag}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches.csv python scratch/prof_rollout.py 128 1 > gpurun_out/${tag}_ncu_launch.log 2>&1
cap() {  # name regex skip
  ncu --set full --import-source on --clock-control none -k regex:$2 -s $3 -c 1 -o gpurun_out/${tag}_$1 -f python scratch/prof_rollout.py 128 1 > gpurun_out/${tag}_ncu_$1.log 2>&1
  ncu -i gpurun_out/${tag}_$1.ncu-rep --page raw --csv > gpurun_out/${tag}_$1.raw.csv
  ncu -i gpurun_out/${tag}_$1.ncu-rep --page source --csv > gpurun_out/${tag}_$1.source.csv
}
cap score_incr k_score_inc 15
cap score_late k_score_inc 36
cap score_step0 k_score_tc 0
cap alpha_incr k_alpha_v3 20
cap alpha_late k_alpha_v3 41
cap alpha_step0 k_alpha_v3 0
cap colblock k_enc_colblock 2
cap ffn k_enc_ffn 2
cap softmax k_softmax_rows 2
cap rowqk k_tc_gemm 2
cap rowpv k_tc_gemm 3
cap merge k_merge 20
ls -la gpurun_out | tail -30
