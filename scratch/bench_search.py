"""Config 3: sampled rollouts per second in Search mode — ONE 50 x 1024 alignment encoded once, S rollouts sampled from the
shared encoder state with on-device Gumbel-max selection (the reference re-encodes for every rollout, finetune_rl_search.py:112).
The second half scores the distinct sampled topologies with the GPU likelihood (GTR+I+G4, neuralnj_b200.likelihood): branch lengths only and with the
model-parameter search, i.e. what RL_Search does per episode."""
import json, sys, time, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/oracle')
from neuralnj_b200 import PhyloATTN, inference_config
from neuralnj_b200.rollout import sample_gumbel
import nnj_oracle as O
S, R, L = 512, 50, 1024
torch.manual_seed(0); m = PhyloATTN(inference_config(), precision="bf16x3").cuda().eval()
data = O.evolved_msa(1, R, L, seed=3).cuda(); mask1 = torch.zeros(1, L, dtype=torch.bool).cuda()
gen = torch.Generator(device="cuda").manual_seed(1)
with torch.no_grad():
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record()
    state = m.encode_zxr(data, mask1).expand(S, -1, -1, -1).contiguous()
    ev[1].record()
    mask = mask1.expand(S, -1).contiguous()
    for _ in range(2):
        g = sample_gumbel((S, R - 1, R * (R - 1) // 2), data.device, gen)
        merges, slp, _ = m.rollout_fused(batch_seq_mask=mask, gumbel=g, state=state)
    torch.cuda.synchronize()
    reps = 5
    ev[2].record()
    for _ in range(reps):
        g = sample_gumbel((S, R - 1, R * (R - 1) // 2), data.device, gen)
        merges, slp, _ = m.rollout_fused(batch_seq_mask=mask, gumbel=g, state=state)
    ev[3].record()
    torch.cuda.synchronize()
ms = ev[2].elapsed_time(ev[3]) / reps
distinct = len({tuple(x.flatten().tolist()) for x in merges.cpu()})
import numpy as np
from neuralnj_b200 import likelihood as LH
mg = merges.cpu().numpy()
uniq = {}
for b in range(S):
    uniq.setdefault(mg[b].tobytes(), b)
idx = list(uniq.values())[:128]
ch = np.stack([LH.children_from_merges(mg[b], R) for b in idx])
masks = LH.onehot_to_masks(data[0].cpu())
labels = [f"t{i}" for i in range(R)]
LH.score_topologies(masks, ch[:4], labels, opt_model=False)          # warm-up (library load, workspace)
torch.cuda.synchronize(); t0 = time.time()
ll_b, _ = LH.score_topologies(masks, ch, labels, model="GTR+I+G", opt_model=False)
torch.cuda.synchronize(); t_b = time.time() - t0
t0 = time.time()
ll_f, _ = LH.score_topologies(masks, ch, labels, model="GTR+I+G", opt_model=True)
torch.cuda.synchronize(); t_f = time.time() - t0
print(json.dumps({"metric": "sampled rollouts/sec, Search mode, one 50 x 1024 alignment, shared encoder pass", "value": round(S / ms * 1e3, 1), "unit": "rollouts/s",
                  "S": S, "ms_per_batch": round(ms, 2), "encode_once_ms": round(ev[0].elapsed_time(ev[1]), 2), "distinct_merge_lists_in_last_batch": distinct,
                  "likelihood_scoring": {"topologies": len(idx), "patterns": int(LH.compress_patterns(masks)[0].shape[1]),
                                         "branch_lengths_only": {"s": round(t_b, 3), "trees_per_s": round(len(idx) / t_b, 1)},
                                         "full_model_optimisation": {"s": round(t_f, 3), "trees_per_s": round(len(idx) / t_f, 1)},
                                         "best_llh": float(ll_f.max()), "mean_gain_from_model_optimisation": float((ll_f - ll_b).mean())}}))
