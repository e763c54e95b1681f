"""Race hunt without a sanitizer: the rollout is bit-reproducible by design (fixed-order reductions), so repeating it many times on
several shapes and demanding identical merges / log-probabilities / logits exposes any ordering bug in the shared-memory exchanges."""
import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/oracle')
from neuralnj_b200 import PhyloATTN, inference_config
import nnj_oracle as O
torch.manual_seed(0); m = PhyloATTN(inference_config(), precision="bf16x3").cuda().eval()
bad = 0
for (B, R, L, reps, logits) in [(128, 50, 1024, 25, False), (64, 20, 256, 40, True), (32, 33, 64, 40, True), (16, 63, 72, 20, True), (40, 12, 128, 40, True)]:
    data = (O.synthetic_msa(B, R, L, seed=7) if R == 50 else O.evolved_msa(B, R, L, seed=7)).cuda()
    mask = torch.zeros(B, L, dtype=torch.bool).cuda()
    ref = None
    for i in range(reps):
        merges, slp, tr = m.rollout_fused(data, mask, want_logits=logits)
        cur = (merges.clone(), slp.clone(), tr.clone() if tr is not None else None)
        if ref is None: ref = cur
        else:
            same = torch.equal(cur[0], ref[0]) and torch.equal(cur[1], ref[1]) and (cur[2] is None or torch.equal(cur[2], ref[2]))
            if not same: bad += 1; print("MISMATCH", B, R, L, "rep", i)
    torch.cuda.synchronize()
    print("shape", B, R, L, "reps", reps, "ok so far" if not bad else "BAD")
print("stress result:", "bit-identical" if not bad else f"{bad} mismatches")
sys.exit(1 if bad else 0)
