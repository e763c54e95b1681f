#!/bin/bash
# usage: bench_kern.sh tag  -> prints bench value and kernel table
tag=$1
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; tail -2 gpurun_out/bench_$tag.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$tag.json").read().strip().splitlines()[-1]); print("$tag", d["value"], d["e2e"]["value"]); print({k:v["ms"] for k,v in d["kernels"].items()})
PY
