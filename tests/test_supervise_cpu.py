"""Row f3 (forward half), CPU side: the label-trajectory sampler against the reference's own (bit-exact, same `random.seed`), and the
oracle restatements - teacher-forced rollout and balanced-ELU loss - against the records the unmodified reference produced
(tests/golden/supervise/, written by `python oracle/make_golden.py supervise`)."""
import os
import random

import numpy as np
import pytest
import torch

from conftest import GOLD

CASES = ["sup_20x256_b2", "sup_50x256", "sup_50x1024"]


def load(name):
    return np.load(os.path.join(GOLD, "supervise", name + ".npz"), allow_pickle=False)


def step_logits(g):
    offs = g["logit_offsets"]
    return [g["logits"][:, offs[t]:offs[t + 1]] for t in range(len(offs) - 1)]


@pytest.mark.parametrize("case", CASES)
def test_trajectory_sampler_equals_reference(case):
    """phydata.load_tree_file + sample_trajectory_set_bottom_top: same actions, same action sets in the same order."""
    from neuralnj_b200.supervise import LabelTree, pair_index
    g = load(case)
    random.seed(int(g["random_seed"][0]))
    R = g["data"].shape[1]
    for b in range(g["actions"].shape[0]):
        acts, sets = LabelTree(str(g["label_newick"][b])).sample_trajectory()
        assert np.array_equal(np.array(acts, dtype=np.int32), g["actions"][b])
        for t in range(R - 2):
            want = [int(x) for x in g["action_set_pair_index"][b, t] if x >= 0]
            assert [pair_index(i, j, R - t) for i, j in sets[t]] == want, (b, t)
        assert sets[-1] == [[0, 1]] and acts[-1] == [0, 1]


def test_label_tree_forms():
    from neuralnj_b200 import NnjError
    from neuralnj_b200.supervise import LabelTree
    a = LabelTree("((taxon3:1,taxon1:1):1,(taxon2:1,taxon4:1):1);")
    b = LabelTree("((taxon1:1,taxon3:1):1,taxon4:1,taxon2:2);")          # trifurcating root: its last two children are grouped (phydata.py:641-646)
    c = LabelTree("((taxon13:1,taxon11:1):1,(taxon12:1,taxon14:1):1);")   # numbering starts at 11
    for t in (a, b, c):
        rng = random.Random(1)
        acts, sets = t.sample_trajectory(rng)
        assert sorted(map(tuple, sets[0])) == [(0, 2), (1, 3)]
        assert acts[-1] == [0, 1] and len(acts) == 3
    # every trajectory is a valid bottom-up order: the action is always a member of the step's set
    big = LabelTree(str(load("sup_50x256")["label_newick"][0]))
    for seed in range(5):
        acts, sets = big.sample_trajectory(random.Random(seed))
        assert all(a in s for a, s in zip(acts, sets))
        assert [len(s) >= 1 for s in sets]
    with pytest.raises(NnjError):
        LabelTree("((taxon1:1,taxon2:1,taxon3:1):1,taxon4:1);")          # not binary below the root
    with pytest.raises(NnjError):
        LabelTree("((taxon1:1,taxon2:1):1,taxon5:1);")                  # indices not contiguous


@pytest.mark.parametrize("case", CASES)
def test_oracle_loss_matches_reference(case):
    import nnj_oracle as O
    g = load(case)
    for k, ep in enumerate(g["loss_epochs"]):
        loss, prec, steps = O.ranking_loss(step_logits(g), g["action_set_pair_index"], int(ep), float(g["ratio_factor"][0]), float(g["margin"][0]))
        assert abs(loss - g["loss"][k]) <= 2e-6 * abs(g["loss"][k])
        assert abs(prec - g["precision"][k]) <= 1e-6
        assert len(steps) == g["data"].shape[1] - 2


def test_oracle_forced_rollout_matches_reference(sd0):
    """train.supervise_rollout(eval=True): the per-step logits along the label trajectory."""
    import nnj_oracle as O
    g = load("sup_20x256_b2")
    r = O.rollout(sd0, torch.from_numpy(g["data"]), torch.from_numpy(g["seq_mask"]), forced_merges=torch.from_numpy(g["actions"]).long())
    for got, want in zip(r["logits"], step_logits(g)):
        want = torch.from_numpy(want)
        assert float((got - want).abs().max() / want.abs().max()) < 1e-5
    assert torch.allclose(r["selected_log_ps"], torch.from_numpy(g["selected_log_ps"]), atol=2e-4)
    assert torch.equal(r["merges"], torch.from_numpy(g["actions"]).long())


def test_action_set_flags_and_guards():
    from neuralnj_b200 import NnjError
    from neuralnj_b200.supervise import action_set_flags, supervise_rollout, topk_ratio, trace_offsets
    sets = [[[[[0, 1], [2, 3]], [[0, 1]], [[0, 1]]]]]                      # one tree of 4 taxa, 3 steps
    f = action_set_flags(sets, 4)
    offs = trace_offsets(4)
    assert offs == [0, 6, 9, 10] and f.shape == (1, 10)
    assert np.flatnonzero(f[0]).tolist() == [0, 5, 6, 9]
    with pytest.raises(NnjError):
        action_set_flags([[[[[0, 4]], [[0, 1]], [[0, 1]]]]], 4)
    assert topk_ratio(0, 0.5) == 0.5 and topk_ratio(30, 0.5) == pytest.approx(0.21875) and topk_ratio(100, 1.0) == 0.25
    with pytest.raises(NotImplementedError):
        supervise_rollout({}, None, None, eval=False)
