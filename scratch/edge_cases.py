import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/oracle')
from neuralnj_b200 import PhyloATTN, inference_config, NnjError
import nnj_oracle as O
torch.manual_seed(0); m = PhyloATTN(inference_config(), precision="bf16x3").cuda().eval()
sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
for (B, R, L) in [(1, 2, 64), (3, 3, 64), (2, 4, 8), (1, 5, 1016), (2, 63, 64), (1, 64, 64), (1, 70, 64), (2, 9, 12), (2, 9, 100)]:
    data = O.evolved_msa(B, R, L, seed=R * 7 + L)
    mask = torch.zeros(B, L, dtype=torch.bool)
    try:
        mg, slp, tr = m.rollout_fused(data.cuda(), mask.cuda(), want_logits=True)
        torch.cuda.synchronize()
        ref = O.rollout(sd, data, mask)
        same = torch.equal(mg.cpu().long(), ref["merges"])
        off = 0; worst = 0.0
        for lg in ref["logits"]:
            p = lg.shape[1]; worst = max(worst, float((tr[:, off:off + p].cpu() - lg).abs().max() / lg.abs().max())); off += p
        print(f"B={B} R={R} L={L}: merges_equal={same} worst_rel_logit_err={worst:.2e}")
    except NnjError as e:
        print(f"B={B} R={R} L={L}: NnjError: {str(e)[:120]}")
