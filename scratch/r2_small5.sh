#!/bin/bash
cd $GRAFT_REPO_ROOT
o=gpurun_out/r2sp; mkdir -p $o
ncu --set full --import-source on --clock-control none -k regex:k_alpha_small -s 18 -c 1 -o $o/asmall -f python scratch/prof_rollout.py 128 1 > $o/ncu_asmall.log 2>&1
ncu -i $o/asmall.ncu-rep --page raw --csv > $o/asmall.raw.csv
ncu -i $o/asmall.ncu-rep --page source --csv > $o/asmall.source.csv
ncu -i $o/asmall.ncu-rep --page details > $o/asmall.details.txt
rm -f $o/asmall.ncu-rep
