"""Stage timeline of k_enc_colblock_tc (CTA 0, thread 0, first 32 work items).  Needs a library built with -DNNJ_COL_TRACE:
   cd neuralnj_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared -cudart static \
        -DNNJ_COL_TRACE *.cu -o ../../scratch/libnnj_trace.so
   NNJ_LIB_PATH=$PWD/scratch/libnnj_trace.so python scratch/col_trace.py [B R L]"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from neuralnj_b200 import PhyloATTN, inference_config, _lib  # noqa: E402
import nnj_oracle as O  # noqa: E402

B, R, L = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (32, 50, 1024)
lib = _lib.lib()
torch.manual_seed(0)
m = PhyloATTN(inference_config(), precision="bf16x3").cuda().eval()
data = O.synthetic_msa(B, R, L, seed=1234).cuda()
mask = torch.zeros(B, L, dtype=torch.bool).cuda()
for _ in range(2):
    m.encode_zxr(data, mask)
torch.cuda.synchronize()
buf = (C.c_longlong * 512)()
rc = lib.nnj_col_trace_read(buf)
assert rc == 0, rc
names = ["load+a_store", "umma1", "resid+LN+a_store", "umma2", "qkv->planes", "attention", "prefetch+publish", "umma3", "store"]
rows = [[buf[i * 16 + k] for k in range(10)] for i in range(32)]
rows = [r for r in rows[2:] if all(v > 0 for v in r)]
print(f"{len(rows)} work items of CTA 0; clocks per stage (mean / min / max):")
tot = 0
for k, nme in enumerate(names):
    d = [r[k + 1] - r[k] for r in rows]
    tot += sum(d) / len(d)
    print(f"  {nme:22s} {sum(d) / len(d):9.0f} {min(d):8d} {max(d):8d}")
gaps = [rows[i + 1][0] - rows[i][9] for i in range(len(rows) - 1)]
print(f"  {'loop turn':22s} {sum(gaps) / len(gaps):9.0f}")
print(f"  total per item {tot + sum(gaps) / len(gaps):.0f} clk")
