#!/bin/bash
cd $GRAFT_REPO_ROOT
o=gpurun_out/r2sp; mkdir -p $o
ncu --set full --import-source on --clock-control none -k regex:k_score_small -s 5 -c 1 -o $o/small16 -f python scratch/prof_rollout.py 128 1 > $o/ncu_small16.log 2>&1
ncu -i $o/small16.ncu-rep --page raw --csv > $o/small16.raw.csv
ncu -i $o/small16.ncu-rep --page source --csv > $o/small16.source.csv
ncu -i $o/small16.ncu-rep --page details > $o/small16.details.txt
rm -f $o/small16.ncu-rep
