#!/bin/bash
cd $GRAFT_REPO_ROOT
o=gpurun_out/r2sp; mkdir -p $o
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tc_paths.py -m gpu -x -q -k "not 200x4096" > $o/pytest.log 2>&1; echo "pytest rc=$?" >> $o/pytest.log
tail -12 $o/pytest.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv -k regex:k_alpha --log-file $o/launches_a.csv python scratch/prof_rollout.py 128 1 > $o/ncu_launch.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/r2sp/launches_a.csv")) if len(r)>5]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value")
for name in ("k_alpha_small","k_alpha_v3"):
    t=[float(r[vi].replace(",",""))/1e3 for r in rows[1:] if name in r[ki]]
    print(name, len(t), round(sum(t)), [round(x) for x in t])
PY
for v in 1 0; do
  NNJ_ALPHA_SMALL=$v timeout 300 python scratch/r2_explore.py 128 50 1024 bf16x3 3 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('alpha_small=$v', d['trees_per_s'], d['classes']['alpha'], d['classes']['pair_score'])"
done
