#!/bin/bash
cd $GRAFT_REPO_ROOT
o=gpurun_out/r2sp; mkdir -p $o
ncu --metrics gpu__time_duration.sum --clock-control none --csv -k regex:k_score --log-file $o/launches.csv python scratch/prof_rollout.py 128 1 > $o/ncu_launch.log 2>&1
cap() {
  ncu --set full --import-source on --clock-control none -k regex:$2 -s $3 -c 1 -o $o/$1 -f python scratch/prof_rollout.py 128 1 > $o/ncu_$1.log 2>&1
  ncu -i $o/$1.ncu-rep --page raw --csv > $o/$1.raw.csv
  ncu -i $o/$1.ncu-rep --page source --csv > $o/$1.source.csv
  rm -f $o/$1.ncu-rep
}
cap small_t2 k_score_small 5
cap small_t1 k_score_small 22
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/r2sp/launches.csv")) if len(r)>5]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value")
for name in ("k_score_small","k_score_inc","k_score_tc"):
    t=[float(r[vi].replace(",",""))/1e3 for r in rows[1:] if name in r[ki]]
    print(name, len(t), round(sum(t)), [round(x) for x in t])
PY
