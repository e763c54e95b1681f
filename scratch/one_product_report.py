"""What plain bf16 operands (ONE product per contraction) would do to parity - numerics only.
Builds nothing itself: run once per library,
    python scratch/one_product_report.py                                      # product build (3-product split)
    NNJ_LIB_PATH=$PWD/scratch/libnnj_1p.so python scratch/one_product_report.py  # experiment build (-DNNJ_ONE_PRODUCT: low parts forced to zero)
and prints one JSON object: per reference-executed golden record (tests/golden) the free-running agreement (identical merges? first differing
step, RF distance of the final tree) and the TEACHER-FORCED agreement (the reference's own trajectory forced through the Gumbel input, so every
step's logits are comparable): largest logit error relative to the step's max |logit|, and how many steps would have picked the reference's pair."""
import json, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import Golden
from neuralnj_b200 import PhyInferEnv, PhyloATTN, inference_config, rf_distance
from neuralnj_b200.environment import format_rtree_topology

CASES = ["ex50x1024_73", "ex50x1024_71", "t20x256_10", "t20x256_103", "t20x256_104", "t20x256_117", "t20x256_120", "t50x256_a", "t100x256_a", "t20x512_a", "t50x512_a"]
cfgs = inference_config()
torch.manual_seed(0)
PREC = os.environ.get("NNJ_REPORT_PRECISION", "bf16x3")     # "bf16": the shipped one-product encoder mode (NNJ_PREC_BF16)
model = PhyloATTN(cfgs, precision=PREC).cuda().eval()
out = {"library": os.environ.get("NNJ_LIB_PATH", "neuralnj_b200/libnnj.so"), "precision": PREC, "cases": {}}
tot_steps = tot_match = 0
for name in CASES:
    g = Golden(name)
    data, mask = g.data.cuda(), g.mask.cuda()
    B, R = data.shape[:2]
    P0 = R * (R - 1) // 2
    ref = g.merges                                   # [B, R-1, 2]
    merges, _, _ = model.rollout_fused(data, mask)
    merges = merges.cpu().long()
    same = bool(torch.equal(merges, ref))
    diff = (merges[0] != ref[0]).any(dim=1).nonzero()
    env = PhyInferEnv(cfgs, torch.device("cuda"))
    seqs = [["A"] * R for _ in range(B)]
    env.init_states(seqs, g.seq_keys, None)
    env.replay_merges(merges)
    rf = rf_distance(env.states[0].subtrees[0].utree_op_str, g.newick[0])
    # teacher forcing: a huge bonus on the reference's pair at every step
    gum = torch.zeros(B, R - 1, P0)
    for b in range(B):
        for t in range(R - 1):
            n = R - t
            i, j = (int(v) for v in ref[b, t])
            gum[b, t, i * n - i * (i + 1) // 2 + (j - i - 1)] = 1e9
    m2, _, trace = model.rollout_fused(data, mask, gumbel=gum.cuda(), want_logits=True)
    assert torch.equal(m2.cpu().long(), ref), "forcing failed"
    trace = trace.cpu()
    off, worst, match, steps = 0, 0.0, 0, 0
    for t, lg in enumerate(g.logits):
        p = lg.shape[1]
        got = trace[:, off:off + p]
        worst = max(worst, float((got - lg).abs().max() / lg.abs().max()))
        if p > 1:
            match += int((got.argmax(1) == lg.argmax(1)).sum()); steps += B
        off += p
    tot_steps += steps; tot_match += match
    out["cases"][name] = {"identical_merges": same, "first_differing_step": int(diff[0]) if len(diff) else None, "rf_to_reference_tree": int(rf),
                          "teacher_forced": {"max_rel_logit_error": worst, "steps_choosing_the_reference_pair": f"{match}/{steps}"}}
out["summary"] = {"records": len(CASES), "identical": sum(c["identical_merges"] for c in out["cases"].values()),
                  "rf_zero": sum(c["rf_to_reference_tree"] == 0 for c in out["cases"].values()),
                  "teacher_forced_step_agreement": round(tot_match / tot_steps, 4),
                  "max_rel_logit_error": max(c["teacher_forced"]["max_rel_logit_error"] for c in out["cases"].values())}
print(json.dumps(out))
