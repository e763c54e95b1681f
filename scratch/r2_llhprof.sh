#!/bin/bash
cd $GRAFT_REPO_ROOT
o=gpurun_out/r2llh; mkdir -p $o
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__cycles_elapsed.max --clock-control none --csv -k regex:k_llh --log-file $o/launches.csv python scratch/llh_prof.py > $o/log1.txt 2>&1
cat $o/launches.csv | tail -12
ncu --set full --import-source on --clock-control none -k regex:k_llh -s 0 -c 1 -o $o/eval -f python scratch/llh_prof.py > $o/log2.txt 2>&1
ncu -i $o/eval.ncu-rep --page raw --csv > $o/eval.raw.csv
ncu -i $o/eval.ncu-rep --page source --csv > $o/eval.source.csv
rm -f $o/eval.ncu-rep
