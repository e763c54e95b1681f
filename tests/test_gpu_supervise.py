"""Row f3 (forward half) on the GPU: the teacher-forced device rollout (NNJ_SELECT_FORCED) and the balanced-ELU loss kernel
(nnj_rank_loss) against the records of the executed reference (tests/golden/supervise/) and against the oracle.

Tolerances: logits as in test_gpu_parity.py (max|delta| / max|logit| per step); the loss kernel on the reference's own logits 1e-5
relative (fp32 elu terms, fp64 sums); the loss end to end 2e-4 relative (the logit error of the rollout carried through
elu(margin + u - s), |d loss| <= max|d logit| * 2)."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from conftest import GOLD
from test_gpu_parity import LOGIT_TOL_BY_PREC

pytestmark = pytest.mark.gpu

CASES = ["sup_20x256_b2", "sup_50x256", "sup_50x1024"]


def load(name):
    return np.load(os.path.join(GOLD, "supervise", name + ".npz"), allow_pickle=False)


def as_batch(g):
    """The reference's collate format (phydata.py:1031-1116) from a golden record; the action sets come back from the stored pair indices."""
    from neuralnj_b200.supervise import LabelTree
    import random
    random.seed(int(g["random_seed"][0]))
    B, R, L = g["data"].shape[:3]
    sets = []
    for b in range(B):
        acts, s = LabelTree(str(g["label_newick"][b])).sample_trajectory()
        assert np.array_equal(np.array(acts), g["actions"][b])
        sets.append([s])
    return {"data": torch.from_numpy(g["data"]), "seq_weights": torch.from_numpy((~g["seq_mask"]).astype(np.float32)),
            "seqs": [["A" * L] * R for _ in range(B)], "seq_keys": [[str(k) for k in row] for row in g["seq_keys"]],
            "actions": torch.from_numpy(g["actions"][:, None]), "actions_set": sets}


def step_logits(g):
    offs = g["logit_offsets"]
    return [torch.from_numpy(g["logits"][:, offs[t]:offs[t + 1]]) for t in range(len(offs) - 1)]


@pytest.mark.parametrize("prec", ["fp32", "bf16x3"])
@pytest.mark.parametrize("case", CASES)
def test_supervise_rollout_matches_reference(case, prec, gpu_models):
    from neuralnj_b200 import PhyInferEnv, inference_config, supervise_rollout, balanced_elu_loss
    g = load(case)
    agent = gpu_models[prec]
    env = PhyInferEnv(inference_config(), torch.device("cuda:0"))
    out = supervise_rollout(as_batch(g), agent, env, eval=True, pretrained=True)
    logitss, sets_list, set_masks, sets, comp_masks, comps, sel = out
    want = step_logits(g)
    R = g["data"].shape[1]
    assert len(out) == 7 and len(logitss) == len(want) == R - 2
    worst = 0.0
    for t, (a, b) in enumerate(zip(logitss, want)):
        worst = max(worst, float((a.cpu() - b).abs().max() / b.abs().max()))
        ref_set = [[int(x) for x in row if x >= 0] for row in g["action_set_pair_index"][:, t]]
        assert sets_list[t] == ref_set                                      # train.py:31-33, the reference's own order
        assert sets[t].shape == set_masks[t].shape and comps[t].shape == comp_masks[t].shape
        P = a.shape[1]
        for bb, row in enumerate(ref_set):
            assert sets[t][bb][set_masks[t][bb]].tolist() == row
            assert comps[t][bb][comp_masks[t][bb]].tolist() == [p for p in range(P) if p not in row]
    assert worst < LOGIT_TOL_BY_PREC[prec], worst
    scale = max(float(w.abs().max()) for w in want)
    assert float((sel.cpu() - torch.from_numpy(g["selected_log_ps"])).abs().max()) < max(1e-3, 2 * LOGIT_TOL_BY_PREC[prec] * scale)
    for k, ep in enumerate(g["loss_epochs"]):
        res = balanced_elu_loss(out, epoch=int(ep), ratio_factor=float(g["ratio_factor"][0]))
        assert abs(res["loss"] - g["loss"][k]) <= 2e-4 * abs(g["loss"][k]) + 4 * LOGIT_TOL_BY_PREC[prec] * scale, (ep, res["loss"], g["loss"][k])
        assert abs(res["precision"] - g["precision"][k]) <= 2e-3
        assert res["step_losses"].shape == (R - 2,)
    # the environment holds the label topology afterwards
    from neuralnj_b200.treeutil import rf_distance
    for b in range(g["data"].shape[0]):
        assert rf_distance(env.states[b].subtrees[0].utree_op_str, str(g["label_newick"][b])) == 0


def _rank_loss(trace, flags, R, margin, ratio):
    from neuralnj_b200 import _lib
    L = _lib.lib()
    B = trace.shape[0]
    ws = torch.empty(L.nnj_rank_loss_workspace_bytes(B, R), dtype=torch.uint8, device="cuda")
    out = torch.empty(R, dtype=torch.float32, device="cuda")
    rc = L.nnj_rank_loss(trace.data_ptr(), flags.data_ptr(), B, R, margin, ratio, out.data_ptr(), ws.data_ptr(), ws.numel(), None)
    assert rc == 0, L.nnj_last_error()
    torch.cuda.synchronize()
    return out.cpu()


@pytest.mark.parametrize("case", CASES)
def test_loss_kernel_on_reference_logits(case):
    """The kernel alone: the reference's logits in, the reference's loss out."""
    from neuralnj_b200.supervise import action_set_flags, topk_ratio, trace_offsets
    import __graft_entry__ as ge
    ge.build()
    g = load(case)
    B, R = g["data"].shape[:2]
    offs = trace_offsets(R)
    trace = torch.zeros(B, offs[-1])
    trace[:, :g["logits"].shape[1]] = torch.from_numpy(g["logits"])
    flags = torch.from_numpy(action_set_flags(as_batch(g)["actions_set"], R))
    for k, ep in enumerate(g["loss_epochs"]):
        out = _rank_loss(trace.cuda(), flags.cuda(), R, 0.5, topk_ratio(int(ep), float(g["ratio_factor"][0])))
        assert abs(float(out[0]) - g["loss"][k]) <= 1e-5 * abs(g["loss"][k]), (float(out[0]), g["loss"][k])
        assert abs(float(out[1]) - g["precision"][k]) <= 1e-6


@pytest.mark.parametrize("B,R,seed", [(3, 7, 0), (5, 33, 1), (2, 130, 2), (1, 200, 3), (300, 5, 4)])
def test_loss_kernel_matches_oracle_random(B, R, seed):
    """Random scores with ties, action sets of different sizes per tree (so the padded width W and the per-tree top-K differ), the
    >48 KB shared-memory path (R = 130, 200) and more trees than threads (B = 300)."""
    import nnj_oracle as O
    from neuralnj_b200.supervise import trace_offsets
    import __graft_entry__ as ge
    ge.build()
    rng = np.random.default_rng(seed)
    offs = trace_offsets(R)
    trace = np.round(rng.normal(0, 3, size=(B, offs[-1])), 1).astype(np.float32)          # one decimal: plenty of equal scores
    flags = np.zeros((B, offs[-1]), dtype=np.uint8)
    width = max(1, R // 2)
    set_index = -np.ones((B, R - 2, width), dtype=np.int64)
    for b in range(B):
        for t in range(R - 2):
            P = offs[t + 1] - offs[t]
            k = int(rng.integers(1, min(width, P - 1) + 1))
            idx = np.sort(rng.choice(P, size=k, replace=False))
            set_index[b, t, :k] = idx
            flags[b, offs[t] + idx] = 1
    steps = [trace[:, offs[t]:offs[t + 1]] for t in range(R - 2)]
    for ratio_factor, epoch in ((0.5, 0), (1.0, 0), (0.5, 60)):
        from neuralnj_b200.supervise import topk_ratio
        want_loss, want_prec, want_steps = O.ranking_loss(steps, set_index, epoch, ratio_factor, 0.5)
        out = _rank_loss(torch.from_numpy(trace).cuda(), torch.from_numpy(flags).cuda(), R, 0.5, topk_ratio(epoch, ratio_factor))
        assert abs(float(out[0]) - want_loss) <= 1e-5 * abs(want_loss)
        assert abs(float(out[1]) - want_prec) <= 1e-6
        assert np.allclose(out[2:].numpy(), np.array(want_steps), rtol=1e-5, atol=1e-6)
    again = _rank_loss(torch.from_numpy(trace).cuda(), torch.from_numpy(flags).cuda(), R, 0.5, topk_ratio(60, 0.5))
    assert torch.equal(again, out)                                                          # fixed summation order


def test_forced_mode_c_abi(gpu_models):
    """nnj_rollout / nnj_rollout_host with NNJ_SELECT_FORCED: follows the given merge lists, equals the argmax rollout when fed the argmax
    trajectory, and falls back to the argmax where an entry is not a pair."""
    from neuralnj_b200 import _lib
    import nnj_oracle as O
    agent = gpu_models["bf16x3"]
    data = O.evolved_msa(3, 9, 64, seed=5)
    mask = torch.zeros(3, 64, dtype=torch.bool)
    m0, slp0, tr0 = agent.rollout_fused(data.cuda(), mask.cuda(), want_logits=True)
    m1, slp1, tr1 = agent.rollout_fused(data.cuda(), mask.cuda(), want_logits=True, forced=m0)
    assert torch.equal(m0, m1) and torch.equal(tr0, tr1) and torch.equal(slp0, slp1)
    # another trajectory: always join the first two nodes
    forced = torch.zeros(3, 8, 2, dtype=torch.int32)
    forced[..., 1] = 1
    m2, _, tr2 = agent.rollout_fused(data.cuda(), mask.cuda(), want_logits=True, forced=forced.cuda())
    assert torch.equal(m2.cpu(), forced)
    ref = O.rollout(O.init_state_dict(0), data, mask, forced_merges=forced.long())
    off = 0
    for lg in ref["logits"]:
        p = lg.shape[1]
        assert float((tr2[:, off:off + p].cpu() - lg).abs().max() / lg.abs().max()) < 4e-5
        off += p
    # invalid entries (i == j, j out of range) -> the argmax of that step, reported in the output
    bad = m0.clone()
    bad[0, 0] = torch.tensor([2, 2]); bad[1, 3] = torch.tensor([0, 99])
    m3, _, _ = agent.rollout_fused(data.cuda(), mask.cuda(), forced=bad)
    assert torch.equal(m3, m0)
    # host-buffer entry point
    L = _lib.lib()
    mh = forced.clone().pin_memory()
    d8 = data.to(torch.int8).contiguous().pin_memory()
    rc = L.nnj_rollout_host(agent.handle(), d8.data_ptr(), None, 3, 9, 64, 2, None, mh.data_ptr(), None)
    assert rc == 0, L.nnj_last_error()
    assert torch.equal(mh, forced)
    rc = L.nnj_rollout_host(agent.handle(), d8.data_ptr(), None, 3, 9, 64, 7, None, mh.data_ptr(), None)
    assert rc != 0 and b"select mode" in L.nnj_last_error()


def test_wrapper_rejects_bad_trajectories(gpu_models):
    from neuralnj_b200 import NnjError, PhyInferEnv, inference_config, supervise_rollout
    g = load("sup_20x256_b2")
    batch = as_batch(g)
    batch["actions"] = batch["actions"].clone()
    batch["actions"][0, 0, 2] = torch.tensor([5, 5])
    env = PhyInferEnv(inference_config(), torch.device("cuda:0"))
    with pytest.raises(NnjError):
        supervise_rollout(batch, gpu_models["fp32"], env, eval=True)


def test_evaluate_on_label_files(gpu_models, tmp_path):
    """supervise.evaluate over (alignment, label tree) files: loss as in the golden record when the same trajectory is drawn."""
    import random
    from neuralnj_b200 import PhyInferEnv, inference_config
    from neuralnj_b200 import supervise as S
    g = load("sup_20x256_b2")
    files = []
    for b, stem in enumerate(("t20x256_10", "t20x256_103")):
        tre = tmp_path / f"{stem}.tre"
        tre.write_text(str(g["label_newick"][b]) + "\n")
        files.append((os.path.join(GOLD, "msa", stem + ".phy"), str(tre)))
    random.seed(int(g["random_seed"][0]))
    env = PhyInferEnv(inference_config(), torch.device("cuda:0"))
    res = S.evaluate(files, gpu_models["bf16x3"], env, cfgs=inference_config(), epoch=0, ratio_factor=0.5)
    assert abs(res["loss"] - g["loss"][0]) <= 1e-3 * abs(g["loss"][0])
    assert 0.0 <= res["argmax_in_action_set"] <= 1.0 and len(res["normalized_rf"]) == 2
    assert all(0.0 <= x <= 1.0 for x in res["normalized_rf"])
    # with the likelihood scorer: the label topology (the tree that generated the data) is not beaten by much, if at all, by an untrained policy's tree
    random.seed(int(g["random_seed"][0]))
    res2 = S.evaluate(files, gpu_models["bf16x3"], env, cfgs=inference_config(), with_likelihood=True)
    assert res2["loss"] == res["loss"] and len(res2["llh_label_tree"]) == 2
    assert all(np.isfinite(res2["llh_label_tree"])) and all(np.isfinite(res2["llh_argmax_tree"]))
    assert all(a > b for a, b in zip(res2["llh_label_tree"], res2["llh_argmax_tree"]))


def test_evaluate_dir_and_cli(gpu_models, tmp_path, capsys):
    """Directory walk (<name>.phy + <name>.tre pairs, one shape per directory) and the command line around it."""
    import json
    import shutil
    from neuralnj_b200 import inference_config
    from neuralnj_b200 import supervise as S
    g = load("sup_20x256_b2")
    d = tmp_path / "len256" / "taxa20"
    d.mkdir(parents=True)
    for b, stem in enumerate(("t20x256_10", "t20x256_103")):
        shutil.copyfile(os.path.join(GOLD, "msa", stem + ".phy"), d / f"{stem}.phy")
        (d / f"{stem}.tre").write_text(str(g["label_newick"][b]) + "\n")
    shutil.copyfile(os.path.join(GOLD, "msa", "t20x256_104.phy"), d / "unlabelled.phy")          # no .tre next to it: skipped
    res = S.evaluate_dir(str(tmp_path), gpu_models["bf16x3"], inference_config(), torch.device("cuda:0"), batch=1, ratio_factor=0.5)
    assert res["files"] == 2 and list(res["by_directory"]) == [os.path.join("len256", "taxa20")]
    assert 0.0 < res["loss"] < 10.0 and 0.0 <= res["normalized_rf_mean"] <= 1.0
    out = S.main(["--data_dir", str(tmp_path), "--limit", "1", "--precision", "fp32"])
    assert out["files"] == 1
    assert json.loads(capsys.readouterr().out)["files"] == 1
    with pytest.raises(S.NnjError):
        S.evaluate_dir(str(tmp_path / "len256" / "taxa20" / "nothing"), gpu_models["fp32"], inference_config(), torch.device("cuda:0"))
