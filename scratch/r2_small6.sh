#!/bin/bash
cd $GRAFT_REPO_ROOT
o=gpurun_out/r2sp; mkdir -p $o
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tc_paths.py -m gpu -x -q -k "not 200x4096" > $o/pytest.log 2>&1; echo "pytest rc=$?" >> $o/pytest.log
tail -4 $o/pytest.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv -k regex:small --log-file $o/launches_s.csv python scratch/prof_rollout.py 128 1 > $o/ncu_launch.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/r2sp/launches_s.csv")) if len(r)>5]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value")
for name in ("k_alpha_small","k_score_small"):
    t=[float(r[vi].replace(",",""))/1e3 for r in rows[1:] if name in r[ki]]
    print(name, len(t), round(sum(t)), [round(x) for x in t])
PY
