#!/bin/bash
cd $GRAFT_REPO_ROOT
o=gpurun_out/r2q; mkdir -p $o
python -c "import __graft_entry__ as g; g.smoke()" > $o/smoke.log 2>&1; echo "smoke rc=$?"; tail -4 $o/smoke.log
timeout 900 python bench.py --steps 3 --warmup 3 > $o/bench.json 2> $o/bench.err; echo "bench rc=$?"; tail -2 $o/bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2q/bench.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d.get("tree_llh"), d.get("latency_b1_ms"))
PY
timeout 600 python scratch/bench_search.py > $o/search.json 2> $o/search.err; cat $o/search.json
