"""Exploration driver (not a bench line): per-kernel-class times of one fused rollout for a given shape.
usage: python scratch/r2_explore.py B R L [precision] [reps]   (NNJ_CHUNK_MAX / NNJ_WS_GB are read by the library)"""
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from neuralnj_b200 import PhyloATTN, inference_config, _lib  # noqa: E402
import nnj_oracle as O  # noqa: E402

B, R, L = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
prec = sys.argv[4] if len(sys.argv) > 4 else "bf16x3"
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
lib = _lib.lib()
torch.manual_seed(0)
m = PhyloATTN(inference_config(), precision=prec).cuda().eval()
data = O.synthetic_msa(B, R, L, seed=1234).cuda()
mask = torch.zeros(B, L, dtype=torch.bool).cuda()
for _ in range(2):
    m.rollout_fused(data, mask)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    m.rollout_fused(data, mask)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
n_cls = lib.nnj_profile_classes()
lib.nnj_profile_enable(1)
m.rollout_fused(data, mask)
cls_ms = (C.c_double * n_cls)()
cls_n = (C.c_int64 * n_cls)()
_lib.check(lib.nnj_profile_read(n_cls, cls_ms, cls_n))
lib.nnj_profile_enable(0)
names = [lib.nnj_profile_name(i).decode() for i in range(n_cls)]
print(json.dumps({"B": B, "R": R, "L": L, "prec": prec, "chunk_max": os.environ.get("NNJ_CHUNK_MAX"), "ms": round(ms, 3),
                  "trees_per_s": round(B / ms * 1e3, 2), "peak_mem_GB": round(torch.cuda.max_memory_allocated() / 1e9, 2),
                  "classes": {n: [round(cls_ms[i], 2), int(cls_n[i])] for i, n in enumerate(names) if cls_n[i]}}))
