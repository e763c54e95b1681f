#!/bin/bash
mkdir -p gpurun_out/r2
cd $GRAFT_REPO_ROOT
tag=${1:-small}
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tc_paths.py -m gpu -x -q -k "not 200x4096" > gpurun_out/r2/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/${tag}_pytest.log
tail -12 gpurun_out/r2/${tag}_pytest.log
for v in 1 0; do
  NNJ_SCORE_SMALL=$v timeout 300 python scratch/r2_explore.py 128 50 1024 bf16x3 3 >> gpurun_out/r2/${tag}.jsonl 2>> gpurun_out/r2/${tag}.err
done
NNJ_SCORE_SMALL=1 timeout 300 python scratch/r2_explore.py 128 20 256 bf16x3 3 >> gpurun_out/r2/${tag}.jsonl 2>> gpurun_out/r2/${tag}.err
NNJ_SCORE_SMALL=0 timeout 300 python scratch/r2_explore.py 128 20 256 bf16x3 3 >> gpurun_out/r2/${tag}.jsonl 2>> gpurun_out/r2/${tag}.err
python - <<'PY'
import json
for l in open("gpurun_out/r2/TAG.jsonl".replace("TAG", "small")):
    d = json.loads(l); print(d["R"], d["L"], d["trees_per_s"], {k: v[0] for k, v in d["classes"].items() if k in ("pair_score", "alpha", "merge")})
PY
tail -5 gpurun_out/r2/${tag}.err
