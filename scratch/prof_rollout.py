import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/oracle')
from neuralnj_b200 import PhyloATTN, inference_config
import nnj_oracle as O
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
prec = sys.argv[3] if len(sys.argv) > 3 else "bf16x3"
torch.manual_seed(0); m = PhyloATTN(inference_config(), precision=prec).cuda().eval()
data = O.synthetic_msa(B, 50, 1024, seed=1234).cuda(); mask = torch.zeros(B, 1024, dtype=torch.bool).cuda()
for _ in range(reps):
    merges, slp, _ = m.rollout_fused(data, mask)
torch.cuda.synchronize()
print("ok", merges[0, :3].tolist())
