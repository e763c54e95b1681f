"""Config 4: 200 taxa x 4096 sites, Argmax.  More than 63 taxa: the NJ loop runs on the fp32 CUDA-core kernels, the encoder's
row attention / FFN on tcgen05 (the fused column block needs <= 128 taxa)."""
import json, sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/oracle')
from neuralnj_b200 import PhyloATTN, inference_config
import nnj_oracle as O
B, R, L = 4, 200, 4096
torch.manual_seed(0); m = PhyloATTN(inference_config(), precision="bf16x3").cuda().eval()
data = O.synthetic_msa(B, R, L, seed=1234).cuda(); mask = torch.zeros(B, L, dtype=torch.bool).cuda()
with torch.no_grad():
    m.rollout_fused(data, mask); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); merges, slp, _ = m.rollout_fused(data, mask); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(json.dumps({"metric": "trees/sec, 200-taxa x 4096-site Argmax inference", "value": round(B / ms * 1e3, 3), "unit": "trees/s", "B": B, "ms": round(ms, 1),
                  "peak_mem_GB": round(torch.cuda.max_memory_allocated() / 1e9, 1)}))
