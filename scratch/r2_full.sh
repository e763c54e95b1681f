#!/bin/bash
mkdir -p gpurun_out/r2
cd $GRAFT_REPO_ROOT
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2/full_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/full_pytest.log
tail -5 gpurun_out/r2/full_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2/full_bench.json 2> gpurun_out/r2/full_bench.err
NNJ_SCORE_SMALL=0 NNJ_ALPHA_SMALL=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2/full_bench_off.json 2> gpurun_out/r2/full_bench_off.err
python - <<'PY'
import json
for f in ("full_bench", "full_bench_off"):
    d = json.load(open(f"gpurun_out/r2/{f}.json"))
    print(f, d["value"], d["e2e"]["value"], d["clocks"]["sm_mhz"], {k: round(v["ms"], 1) for k, v in d["kernels"].items() if k in ("alpha", "pair_score", "merge", "col_attn")})
PY
