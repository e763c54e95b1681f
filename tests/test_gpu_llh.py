"""GPU tests of the tree-likelihood path (csrc/nnj_llh.cu through the C ABI) against the CPU oracle (oracle/llh_oracle.py).
fp64 on both sides: log-likelihoods agree to 1e-9 relative; optimised branch lengths to 1e-6."""
import os
import sys
import time

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import llh_oracle as O  # noqa: E402

from neuralnj_b200 import likelihood as LH  # noqa: E402
from neuralnj_b200.treeutil import rf_distance  # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = os.path.join(ROOT, "tests", "golden")


def _case(R, L, seed, gap=0.05):
    rng = np.random.default_rng(seed)
    gen = O.Model(rates6=(1.2, 3.1, 0.8, 1.1, 4.2, 1.0), freqs=(0.3, 0.2, 0.2, 0.3), alpha=0.7, pinv=0.15)
    ch, bl = O.random_tree(R, rng)
    tips = O.simulate(ch, bl, R, L, gen, rng, gap_frac=gap)
    return ch, bl, tips, rng


def _sub_model(B, rates, freqs, alpha, pinv):
    sm = LH.SubstModel("GTR+I+G", np.asarray(freqs), B)
    sm.rates[:] = rates
    sm.alpha[:] = alpha
    sm.pinv[:] = pinv
    return sm


@pytest.mark.parametrize("R,L", [(3, 17), (5, 64), (12, 300), (50, 1024), (130, 700)])
def test_loglik_matches_oracle(R, L):
    ch, bl, tips, rng = _case(R, L, seed=R)
    tips[0, :2] = 5
    pat, w = O.compress_patterns(tips)
    rates, freqs = (0.7, 2.0, 1.3, 0.9, 3.0, 1.0), (0.22, 0.28, 0.31, 0.19)
    for alpha, pinv in ((0.5, 0.0), (1.7, 0.3)):
        ref = O.loglik(ch, bl, pat, w, O.Model(rates, freqs, alpha, pinv))
        eng = LH.TreeLikelihood(pat, w)
        got = eng.loglik(ch[None], bl[None], _sub_model(1, rates, freqs, alpha, pinv))
        assert abs(got[0] - ref) < 1e-9 * abs(ref)


def test_batch_of_trees_and_alignments():
    """B trees in one launch: different topologies, branch lengths, models AND alignments; each equals its own B = 1 result."""
    B, R, L = 7, 9, 120
    cases = [_case(R, L, seed=100 + b) for b in range(B)]
    ch = np.stack([c[0] for c in cases]); bl = np.stack([c[1] for c in cases]); tips = np.stack([c[2] for c in cases])
    sm = LH.SubstModel("GTR+I+G", np.full(4, 0.25), B)
    rng = np.random.default_rng(0)
    sm.rates[:, :5] = rng.uniform(0.3, 4.0, size=(B, 5))
    sm.freqs = rng.dirichlet(np.full(4, 20.0), size=B)
    sm.alpha = rng.uniform(0.2, 3.0, size=B)
    sm.pinv = rng.uniform(0.0, 0.4, size=B)
    got = LH.TreeLikelihood(tips).loglik(ch, bl, sm)
    for b in range(B):
        ref = O.loglik(ch[b], bl[b], tips[b], np.ones(L), O.Model(sm.rates[b], sm.freqs[b], sm.alpha[b], sm.pinv[b]))
        assert abs(got[b] - ref) < 1e-9 * abs(ref)
    again = LH.TreeLikelihood(tips).loglik(ch, bl, sm)
    assert np.array_equal(got, again)                      # fixed-order reductions: bit-reproducible


@pytest.mark.parametrize("R,L", [(4, 80), (12, 300), (30, 500)])
def test_branch_optimiser_matches_oracle(R, L):
    ch, bl, tips, rng = _case(R, L, seed=7 * R)
    pat, w = O.compress_patterns(tips)
    rates, freqs, alpha, pinv = (0.7, 2.0, 1.3, 0.9, 3.0, 1.0), (0.22, 0.28, 0.31, 0.19), 0.8, 0.1
    start = np.full(2 * R - 2, LH.BRLEN_DEFAULT)
    t_ref, b_ref, a_ref = O.optimize_branches(ch, start, pat, w, O.Model(rates, freqs, alpha, pinv))
    eng = LH.TreeLikelihood(pat, w)
    sm = _sub_model(1, rates, freqs, alpha, pinv)
    t, before, after = eng.optimize_branches(ch[None], start[None], sm)
    assert abs(before[0] - b_ref) < 1e-9 * abs(b_ref) and abs(after[0] - a_ref) < 1e-8 * abs(a_ref)
    assert after[0] > before[0]
    assert np.allclose(t[0], t_ref, rtol=1e-5, atol=1e-8)
    assert abs(eng.loglik(ch[None], t, sm)[0] - after[0]) < 1e-8 * abs(after[0])        # the returned lengths reproduce the reported maximum


def test_full_optimisation_matches_oracle_and_recovers_the_model():
    R, L = 10, 1500
    ch, bl, tips, rng = _case(R, L, seed=5, gap=0.02)
    pat, w = O.compress_patterns(tips)
    freqs = LH.empirical_freqs(tips)
    om = O.Model(freqs=freqs, alpha=1.0, pinv=0.0)
    t_ref, b_ref, a_ref = O.optimize_all(ch, np.full(2 * R - 2, 0.1), pat, w, om)
    sm = LH.SubstModel("GTR+I+G", freqs, 1)
    t, before, after = LH.TreeLikelihood(pat, w).optimize_all(ch[None], np.full((1, 2 * R - 2), 0.1), sm)
    assert abs(after[0] - a_ref) < 1e-5 * abs(a_ref)
    assert np.allclose(sm.rates[0], om.rates6, rtol=2e-3) and abs(sm.alpha[0] - om.alpha) < 2e-3 * om.alpha and abs(sm.pinv[0] - om.pinv) < 2e-3
    # the generating model was AC 1.2, AG 3.1, AT 0.8, CG 1.1, CT 4.2, alpha 0.7, p_inv 0.15: the estimate lands near it
    assert 2.0 < sm.rates[0, 1] < 4.5 and 2.8 < sm.rates[0, 4] < 6.0 and 0.3 < sm.alpha[0] < 2.0
    assert after[0] >= O.loglik(ch, bl, pat, w, O.Model((1.2, 3.1, 0.8, 1.1, 4.2, 1.0), (0.3, 0.2, 0.2, 0.3), 0.7, 0.15)) - 1e-6


def test_raxmlpy_style_entry_points():
    """optimize_brlen / compute_llh with the reference's signatures (RAxMLpy/raxmlpy/core.py:6-12, called as in environment.py:365-379)."""
    R, L = 8, 400
    ch, bl, tips, rng = _case(R, L, seed=21, gap=0.0)
    labels = [f"taxon{i}" for i in range(R)]
    seqs = ["".join("ACGT"[int(np.log2(m))] for m in row) for row in tips]
    msa = {"labels": labels, "sequences": seqs}
    rooted = LH.tuples_to_newick(LH.tuples_with_lengths(ch, bl, labels, unrooted=False))
    jc = LH.compute_llh(rooted, msa, is_root=True, model="JC", opt_model=False)
    assert abs(jc - O.loglik(ch, bl, tips, np.ones(L), O.Model(gamma=False))) < 1e-5          # Newick carries 8 decimals
    fixed = LH.compute_llh(rooted, msa, is_root=True, model="GTR+I+G", opt_model=False)
    tuned = LH.compute_llh(rooted, msa, is_root=True, model="GTR+I+G", opt_model=True)          # model parameters only, branch lengths as given
    assert tuned > fixed
    newick, before, after = LH.optimize_brlen(rooted, msa, is_root=True, iters=3, model="GTR+I+G", opt_model=True)
    assert after >= tuned - 1.0 and abs(before - fixed) < 1e-6 * abs(fixed)
    assert after > before and rf_distance(newick, rooted) == 0 and newick.count(":") == 2 * R - 3 + 0 * R
    assert all(lbl in newick for lbl in labels)
    # a wrong topology scores worse than the generating one after optimisation
    perm = labels[::-1]
    wrong = LH.tuples_to_newick(LH.tuples_with_lengths(ch, bl, perm, unrooted=False))
    if rf_distance(wrong, rooted) > 0:
        _, _, after_wrong = LH.optimize_brlen(wrong, msa, is_root=True, model="GTR+I+G", opt_model=True)
        assert after_wrong < after


def test_env_branch_optimize_and_search_scoring(golden):
    """branch_optimize=True (environment.py:648-671) and RL_Search's default scorer run on the GPU likelihood: finite scores,
    optimised lengths on the trees, the best tree is the highest-likelihood topology among those sampled."""
    from neuralnj_b200 import PhyInferEnv, PhyloATTN, RL_Search, inference_config, load_pi_instance, reinforce_rollout
    cfgs = inference_config()
    path = os.path.join(GOLD, "msa", "t20x256_120.phy")
    torch.manual_seed(0)
    model = PhyloATTN(cfgs, precision="bf16x3").to("cuda:0").eval()
    env = PhyInferEnv(cfgs, torch.device("cuda:0"))
    batch = load_pi_instance(path)
    _, _, scores, best = reinforce_rollout(batch, model, env, cfgs, eval=True, argmax=True, branch_optimize=True)
    assert scores.shape == (1,) and np.isfinite(float(scores[0])) and float(scores[0]) < 0
    assert rf_distance(best, golden("t20x256_120").newick[0]) == 0 and "0.12345" not in best
    _, _, scores_step, best_step = reinforce_rollout(batch, model, env, cfgs, eval=True, argmax=True, branch_optimize=True, fused=False)
    assert abs(float(scores_step[0]) - float(scores[0])) < 1e-6 * abs(float(scores[0])) and best_step == best
    # the score is the likelihood of that tree: re-score the returned Newick through the raxmlpy-style entry point
    msa = {"labels": batch["seq_keys"][0], "sequences": batch["seqs"][0]}
    _, _, again = LH.optimize_brlen(best, msa, model="GTR+I+G", opt_model=True)
    assert abs(again - float(scores[0])) < 1.5          # the outer loop stops when a round gains < 1 log-unit (lh_epsilon of the reference)
    cfgs.env.batch_size = 6
    cfgs.num_episodes = 2
    gen = torch.Generator(device="cuda:0").manual_seed(3)
    out = RL_Search(cfgs, path, model, env, stop_step=2, generator=gen)
    assert np.isfinite(out["the_best_score"]) and out["the_best_score"] < 0 and out["distinct_topologies"] >= 1
    assert all(k in out["the_best_tree"] for k in batch["seq_keys"][0])
    _, _, chk = LH.optimize_brlen(out["the_best_tree"], msa, model="GTR+I+G", opt_model=True)
    assert abs(chk - out["the_best_score"]) < 1.5


def test_scoring_throughput_50x1024(capsys):
    """Config-3 shape: 64 candidate topologies of one 50 x 1024 alignment, branch lengths + model optimised together."""
    R, L, B = 50, 1024, 64
    ch0, bl, tips, rng = _case(R, L, seed=9)
    ch = np.stack([O.random_tree(R, rng)[0] for _ in range(B - 1)] + [ch0])
    labels = [f"t{i}" for i in range(R)]
    torch.cuda.synchronize()
    t0 = time.time()
    ll, brl = LH.score_topologies(tips, ch, labels, model="GTR+I+G", opt_model=True)
    dt = time.time() - t0
    assert np.isfinite(ll).all() and int(np.argmax(ll)) == B - 1             # the generating topology wins
    t1 = time.time()
    ll2, _ = LH.score_topologies(tips, ch, labels, model="GTR+I+G", opt_model=False)
    dt2 = time.time() - t1
    assert (ll >= ll2 - 1e-6).all()
    with capsys.disabled():
        print(f"\n[llh] {B} topologies x 50 taxa x 1024 sites: full optimisation {dt:.2f} s ({B / dt:.1f} trees/s), branch lengths only {dt2:.2f} s ({B / dt2:.1f} trees/s)")
