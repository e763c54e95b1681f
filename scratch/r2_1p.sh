#!/bin/bash
cd $GRAFT_REPO_ROOT
o=gpurun_out/r2_1p; mkdir -p $o
timeout 250 python scratch/one_product_report.py > $o/x3.json 2> $o/x3.err; echo "x3 rc=$?"; tail -2 $o/x3.err
NNJ_LIB_PATH=$GRAFT_REPO_ROOT/scratch/libnnj_1p.so timeout 250 python scratch/one_product_report.py > $o/x1.json 2> $o/x1.err; echo "x1 rc=$?"; tail -2 $o/x1.err
python - <<'PY'
import json
for f in ("x3", "x1"):
    try:
        d = json.loads(open(f"gpurun_out/r2_1p/{f}.json").read().strip().splitlines()[-1]); print(f, d["summary"])
    except Exception as e: print(f, "failed", e)
PY
