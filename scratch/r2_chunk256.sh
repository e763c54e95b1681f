#!/bin/bash
cd $GRAFT_REPO_ROOT
o=gpurun_out/r2c; mkdir -p $o
for rep in 1 2; do
timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras > $o/def_$rep.json 2> $o/def_$rep.err
NNJ_WS_GB=100 NNJ_CHUNK_MAX=256 timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras > $o/c256_$rep.json 2> $o/c256_$rep.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2c/*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["value"], d["e2e"]["value"], d["clocks"]["sm_mhz"])
    except Exception as e: print(f, "fail", e)
PY
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
