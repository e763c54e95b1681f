import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/oracle')
from neuralnj_b200 import PhyloATTN, inference_config
import nnj_oracle as O
R = int(sys.argv[1]) if len(sys.argv) > 1 else 20
L = int(sys.argv[2]) if len(sys.argv) > 2 else 256
B = int(sys.argv[3]) if len(sys.argv) > 3 else 1
torch.manual_seed(0); m32 = PhyloATTN(inference_config(), precision="fp32").cuda().eval()
torch.manual_seed(0); mtc = PhyloATTN(inference_config(), precision="bf16x3").cuda().eval()
data = O.evolved_msa(B, R, L, seed=5).cuda(); mask = torch.zeros(B, L, dtype=torch.bool).cuda()
m1, _, t1 = m32.rollout_fused(data, mask, want_logits=True)
m2, _, t2 = mtc.rollout_fused(data, mask, want_logits=True)
torch.cuda.synchronize()
off = 0
for step, n in enumerate(range(R, 1, -1)):
    p = n * (n - 1) // 2
    a, b = t1[:, off:off + p], t2[:, off:off + p]
    same = bool((m1[:, step] == m2[:, step]).all())
    print(f"step {step:2d} n={n:3d} max|dlogit|={float((a - b).abs().max()):.3e} scale={float(a.abs().max()):.2f} merge_same={same}")
    off += p
    if not same: break
