import sys, time, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests'); sys.path.insert(0, '/root/repo/oracle')
import nnj_oracle as O
from conftest import Golden, ALL_CASES
from neuralnj_b200 import PhyloATTN, inference_config
torch.manual_seed(0); m32 = PhyloATTN(inference_config()).cuda().eval()
torch.manual_seed(0); mtc = PhyloATTN(inference_config(), precision="bf16x3").cuda().eval()
sd = O.init_state_dict(0)
g = Golden("t20x256_10")
want = O.encode(sd, g.data, g.mask)
for name, m in (("fp32", m32), ("bf16x3", mtc)):
    got = m.encode_zxr(g.data.cuda(), g.mask.cuda()).cpu()
    print(name, "encoder max abs err", float((got - want).abs().max()), "scale", float(want.abs().max()))
for c in ALL_CASES:
    g = Golden(c)
    merges, slp, trace = mtc.rollout_fused(g.data.cuda(), g.mask.cuda(), want_logits=True)
    same = torch.equal(merges.cpu().long(), g.merges)
    tr = trace.cpu(); off = 0; worst = 0
    for lg in g.logits:
        p = lg.shape[1]; worst = max(worst, float((tr[:, off:off+p] - lg).abs().max() / lg.abs().max())); off += p
    first = int((merges.cpu().long() != g.merges).any(-1).any(0).nonzero()[0]) if not same else -1
    print(c, "merges identical:", same, "first diff step", first, "max rel dlogit %.2e" % worst)
