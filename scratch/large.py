import sys, time, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/oracle')
from neuralnj_b200 import PhyloATTN, inference_config
import nnj_oracle as O
R, L, B = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
for prec in ("bf16x3", "fp32"):
    torch.manual_seed(0); m = PhyloATTN(inference_config(), precision=prec).cuda().eval()
    data = O.synthetic_msa(B, R, L, seed=7).cuda(); mask = torch.zeros(B, L, dtype=torch.bool).cuda()
    torch.cuda.synchronize(); t0 = time.time()
    mg, slp, _ = m.rollout_fused(data, mask)
    torch.cuda.synchronize(); t1 = time.time()
    mg2, _, _ = m.rollout_fused(data, mask)
    torch.cuda.synchronize(); t2 = time.time()
    ok = all(0 <= int(mg[b, t, 0]) < int(mg[b, t, 1]) < R - t for b in range(B) for t in range(R - 1))
    print(prec, f"{R}x{L} B={B}: first {t1 - t0:.2f}s second {t2 - t1:.2f}s valid={ok} deterministic={torch.equal(mg, mg2)} finite={bool(torch.isfinite(slp).all())} mem={torch.cuda.max_memory_allocated() / 2**30:.1f}GiB")
