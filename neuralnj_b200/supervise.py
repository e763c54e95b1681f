"""Teacher-forced rollouts and the pre-training loss - the FORWARD half of the reference's training step.

What the reference does per training step (`train.py:433-554`): sample a bottom-up trajectory of the label tree together with, for
every step, the set of all pairs that are cherries of the label tree at that point (`phydata.py:779-834`); run the rollout with the
actions forced to that trajectory (`supervise_rollout`, `train.py:43-161`); rank the action-set scores of every step against the K
best other scores with a margin (balanced-ELU loss, `train.py:448-545`); back-propagate.  Here the forced rollout runs on the fused
device loop (`nnj_rollout` with `NNJ_SELECT_FORCED`) and the loss on `nnj_rank_loss`, which is what validation
(`supervise_rollout(eval=True)`, `train.py:356-430`) and loss monitoring of a checkpoint need.  The backward pass and the optimiser
are not implemented: `supervise_rollout(eval=False)` raises.
"""
from __future__ import annotations

import math
import random as _random
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import NnjError, check
from .treeutil import normalized_rf, treestr_to_tuples

MARGIN = 0.5            # train.py:478


# ------------------------------------------------------------------------------------------------------------------ label trees
class LabelTree:
    """The label tree as `phydata.load_tree_file` leaves it (`phydata.py:636-682`): rooted and binary (a trifurcating root is
    resolved by grouping its last two children), every node keyed by the smallest taxon index below it, children ordered by that key.

    Arrays over node ids (leaves first, then inner nodes in post-order, the root last): `key`, `parent`, `left`, `right` (-1 for leaves)."""

    def __init__(self, newick: str, pos: int = 5):
        tup = treestr_to_tuples(newick)
        nested = self._strip(tup)
        if not isinstance(nested, tuple):
            raise NnjError("label tree: a single leaf is not a tree")
        if len(nested) == 3:                                    # clades3to2, phydata.py:641-646
            nested = (nested[0], (nested[2], nested[1]))
        self.key: List[int] = []
        self.left: List[int] = []
        self.right: List[int] = []
        self.names: List[Optional[str]] = []
        leaves: List[Tuple[int, str]] = []

        def collect(t):
            if isinstance(t, str):
                leaves.append((int(t[pos:]) - 1, t))
            else:
                if len(t) != 2:
                    raise NnjError("label tree: not binary below the root")
                collect(t[0]); collect(t[1])

        collect(nested)
        leaves.sort()
        base = leaves[0][0]                                     # taxon numbering may start anywhere (phydata.py:240-242)
        self.n_leaves = len(leaves)
        self.leaf_id = {name: k - base for k, name in leaves}
        if sorted(self.leaf_id.values()) != list(range(self.n_leaves)):
            raise NnjError("label tree: taxon indices are not contiguous")
        for k, name in leaves:
            self.key.append(k - base); self.left.append(-1); self.right.append(-1); self.names.append(name)

        def build(t) -> int:
            if isinstance(t, str):
                return self.leaf_id[t]
            a, b = build(t[0]), build(t[1])
            if self.key[a] > self.key[b]:                      # sorted_bio_tree, phydata.py:664-679
                a, b = b, a
            self.key.append(self.key[a]); self.left.append(a); self.right.append(b); self.names.append(None)
            return len(self.key) - 1

        self.root = build(nested)
        self.parent = [-1] * len(self.key)
        for v in range(self.n_leaves, len(self.key)):
            self.parent[self.left[v]] = v
            self.parent[self.right[v]] = v

    @staticmethod
    def _strip(t):
        """Drop the branch lengths of `treestr_to_tuples`' (child, length, child, length ...) form."""
        if isinstance(t, str):
            return t
        return tuple(LabelTree._strip(x) for x in t if isinstance(x, (tuple, str)))

    def sample_trajectory(self, rng=_random) -> Tuple[List[List[int]], List[List[List[int]]]]:
        """`sample_trajectory_set_bottom_top` (`phydata.py:779-834`): a uniformly random bottom-up join order of the label tree.

        Returns (actions, action_sets): per step the chosen pair [i, j] and the list of all admissible pairs - positions in the
        current node list, which is ordered by smallest taxon index (the order `PhyInferEnv.step` keeps).  Consumes `rng.choice`
        exactly as the reference does, so `random.seed(s)` reproduces the reference's trajectory."""
        live = list(range(self.n_leaves))
        actions, action_sets = [], []
        for _ in range(self.n_leaves - 1):
            live.sort(key=lambda v: self.key[v])
            where = {v: i for i, v in enumerate(live)}
            seen = {}
            action_set = []
            for v in live:                                      # a cherry is reported when its second child is met
                m = self.parent[v]
                seen[m] = seen.get(m, 0) + 1
                if seen[m] == 2:
                    action_set.append([where[self.left[m]], where[self.right[m]]])
            act = rng.choice(action_set)
            i, j = act
            mom = self.parent[live[i]]
            live.pop(j)
            live[i] = mom
            actions.append(act)
            action_sets.append(action_set)
        return actions, action_sets


def load_tree_file(path: str, pos: int = 5) -> LabelTree:
    with open(path) as f:
        return LabelTree(f.read().strip(), pos=pos)


def pair_index(i: int, j: int, n: int) -> int:
    """Position of (i, j), i < j, in `itertools.combinations(range(n), 2)` (`environment.py:455-462`)."""
    return i * n - i * (i + 1) // 2 + (j - i - 1)


def trace_offsets(R: int) -> List[int]:
    offs = [0]
    for n in range(R, 1, -1):
        offs.append(offs[-1] + n * (n - 1) // 2)
    return offs


def action_set_flags(batch_action_set, R: int) -> np.ndarray:
    """uint8 [B, trace_len] in the layout of the rollout's logits trace: 1 where the pair belongs to the step's action set
    (`prepare_action_sets`, `train.py:30-39`).  `batch_action_set[b][0][t]` = list of [i, j] (the collate format, `phydata.py:1031-1116`)."""
    offs = trace_offsets(R)
    flags = np.zeros((len(batch_action_set), offs[-1]), dtype=np.uint8)
    for b, per_traj in enumerate(batch_action_set):
        steps = per_traj[0]
        if len(steps) != R - 1:
            raise NnjError(f"action sets: {len(steps)} steps for {R} taxa")
        for t, pairs in enumerate(steps):
            n = R - t
            for i, j in pairs:
                if not (0 <= i < j < n):
                    raise NnjError(f"action sets: ({i}, {j}) is not a pair of {n} nodes (tree {b}, step {t})")
                flags[b, offs[t] + pair_index(i, j, n)] = 1
    return flags


# ------------------------------------------------------------------------------------------------------------------ the rollout
class SupervisedRollout(tuple):
    """The 7-tuple `supervise_rollout` returns in the reference (`train.py:160-161`) - logitss, actions_sets_list, actions_set_masks,
    actions_sets, actions_set_complement_masks, actions_sets_complement, selected_log_ps - plus the device buffers the loss kernel
    reads: `.trace` fp32 [B, trace_len], `.in_set` uint8 [B, trace_len], `.taxa`."""
    trace: torch.Tensor
    in_set: torch.Tensor
    taxa: int


def _pad_index_lists(rows: Sequence[np.ndarray], device) -> Tuple[torch.Tensor, torch.Tensor]:
    """`utils.pad_array_mask` (`utils.py:198-209`): zero-padded int64 [B, width] and its bool mask."""
    width = max(len(r) for r in rows)
    idx = np.zeros((len(rows), width), dtype=np.int64)
    mask = np.zeros((len(rows), width), dtype=bool)
    for b, r in enumerate(rows):
        idx[b, :len(r)] = r
        mask[b, :len(r)] = True
    return torch.from_numpy(idx).to(device), torch.from_numpy(mask).to(device)


def supervise_rollout(batch, agent, env, eval=False, pretrained=True, action_set=True, branch_optimize=False):
    """`train.supervise_rollout` (`train.py:43-161`), forward only: the rollout follows `batch['actions'][:, 0]`.

    `batch` is the reference's collate format: 'data' int8 [B,R,L,4], 'seq_weights' [B,L], 'seqs', 'seq_keys',
    'actions' int [B, trajectories, R-1, 2], 'actions_set' (list per tree of list per trajectory of list per step of [i, j]).
    Returns a `SupervisedRollout`; with `branch_optimize=True` the reference's two extra values (scores, best_tree) follow as
    attributes `.scores`, `.best_tree` (the tuple itself keeps seven entries)."""
    if not eval:
        raise NotImplementedError("supervise_rollout(eval=False) needs the backward pass of the encoder and of the NJ loop "
                                  "(train.py:547): not part of this library - see DESIGN.md section 8")
    if not pretrained:
        raise NotImplementedError("supervise_rollout(pretrained=False) leaves `actions` undefined in the reference as well (train.py:116)")
    device = next(agent.parameters()).device
    data = batch["data"].to(device)
    B, R = data.shape[:2]
    seq_mask = batch["seq_weights"].to(device) == 0
    forced = torch.as_tensor(np.asarray(batch["actions"])[:, 0], dtype=torch.int32)
    if tuple(forced.shape) != (B, R - 1, 2):
        raise NnjError(f"supervise_rollout: actions must be [B, trajectories, {R - 1}, 2]")
    fm = forced.numpy()
    n_live = R - np.arange(R - 1)
    if not ((fm[..., 0] >= 0) & (fm[..., 0] < fm[..., 1]) & (fm[..., 1] < n_live[None, :])).all():
        raise NnjError("supervise_rollout: an action is not a pair i < j of the current node list")
    env.init_states(batch["seqs"], batch["seq_keys"], data)
    agent.eval()
    with torch.no_grad():
        merges, slp, trace = agent.rollout_fused(data, seq_mask, want_logits=True, forced=forced.to(device))
    if not torch.equal(merges.cpu(), forced):
        raise NnjError("supervise_rollout: the device loop did not follow the forced trajectory")
    offs = trace_offsets(R)
    steps = R - 2                                   # the last step has one candidate and is not recorded (train.py:131-136)
    logitss = [trace[:, offs[t]:offs[t + 1]] for t in range(steps)]
    selected_log_ps = slp[:, :steps]
    sets_list, set_masks, sets, comp_masks, comps = [], [], [], [], []
    in_set = None
    if action_set:
        flags = action_set_flags(batch["actions_set"], R)
        in_set = torch.from_numpy(flags).to(device)
        for t in range(steps):
            f = flags[:, offs[t]:offs[t + 1]]
            n = R - t                                # the set in the batch's own order (train.py:31-33), the complement ascending (:34)
            rows = [np.array([pair_index(i, j, n) for i, j in batch["actions_set"][b][0][t]], dtype=np.int64) for b in range(B)]
            crow = [np.flatnonzero(f[b] == 0) for b in range(B)]
            sets_list.append([r.tolist() for r in rows])
            idx, mask = _pad_index_lists(rows, device)
            sets.append(idx); set_masks.append(mask)
            idx, mask = _pad_index_lists(crow, device)
            comps.append(idx); comp_masks.append(mask)
    env.replay_merges(merges, branch_optimize=branch_optimize)
    out = SupervisedRollout((logitss, sets_list, set_masks, sets, comp_masks, comps, selected_log_ps))
    out.trace, out.in_set, out.taxa = trace, in_set, R
    if branch_optimize:
        out.scores, _, _, out.best_tree = env.evaluate_loglikelihood()
    return out


# ------------------------------------------------------------------------------------------------------------------ the loss
def topk_ratio(epoch: int, ratio_factor: float) -> float:
    """`train.py:495`: the share of the complement that enters the loss shrinks from 1 to 1/4 over the epochs."""
    return max(1 - 3 / (4 * 20) * epoch / 2, 1 / 4) * ratio_factor


def balanced_elu_loss(rollout: SupervisedRollout, epoch: int = 0, ratio_factor: float = 0.5, margin: float = MARGIN) -> dict:
    """The BALANCED_ELU_LOSS branch of the training step (`train.py:448-545`) on the device (`nnj_rank_loss`).

    Returns {'loss': policy_loss, 'precision': ..., 'step_losses': fp32 [R-2] tensor}.  `ratio_factor` is `cfgs.ratio_factor`
    (0.5 in config/pretrain_mix.yaml)."""
    if rollout.in_set is None:
        raise NnjError("balanced_elu_loss: the rollout was made with action_set=False")
    trace, flags, R = rollout.trace.contiguous(), rollout.in_set.contiguous(), rollout.taxa
    B = trace.shape[0]
    if R < 3:
        raise NnjError("balanced_elu_loss: needs at least 3 taxa")
    L = _lib.lib()
    nbytes = L.nnj_rank_loss_workspace_bytes(B, R)
    if nbytes < 0:
        raise NnjError("balanced_elu_loss: unsupported shape (3 <= taxa <= 256)")
    ws = torch.empty(nbytes, dtype=torch.uint8, device=trace.device)
    out = torch.empty(R, dtype=torch.float32, device=trace.device)
    with torch.cuda.device(trace.device):
        check(L.nnj_rank_loss(trace.data_ptr(), flags.data_ptr(), B, R, float(margin), float(topk_ratio(epoch, ratio_factor)),
                              out.data_ptr(), ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream), "nnj_rank_loss")
    host = out.cpu()
    return {"loss": float(host[0]), "precision": float(host[1]), "step_losses": host[2:]}


# ------------------------------------------------------------------------------------------------------------------ validation
def make_batch(files: Sequence[Tuple[str, str]], rng=_random) -> dict:
    """Collate (alignment file, label tree file) pairs of one shape into the batch format above, one sampled trajectory each."""
    from .phydata import load_pi_instance
    items = []
    for phy, tre in files:
        b = load_pi_instance(phy)
        tree = load_tree_file(tre)
        if tree.n_leaves != b["data"].shape[1]:
            raise NnjError(f"{tre}: {tree.n_leaves} leaves for {b['data'].shape[1]} sequences")
        acts, sets = tree.sample_trajectory(rng)
        with open(tre) as f:
            items.append((b, tree, acts, sets, f.read().strip()))
    return {
        "data": torch.cat([it[0]["data"] for it in items]),
        "seq_weights": torch.cat([it[0]["seq_weights"] for it in items]),
        "seqs": [it[0]["seqs"][0] for it in items],
        "seq_keys": [it[0]["seq_keys"][0] for it in items],
        "trees": [it[1] for it in items],
        "actions": torch.from_numpy(np.array([[it[2]] for it in items], dtype=np.int32)),
        "actions_set": [[it[3]] for it in items],
        "label_newick": [it[4] for it in items],
        "file_paths": [f[0] for f in files],
    }


def evaluate(files: Sequence[Tuple[str, str]], agent, env, cfgs=None, epoch: int = 0, ratio_factor: float = 0.5, rng=_random,
             with_likelihood: bool = False) -> dict:
    """The quantities `eval_on_dataset` logs for a validation set (`train.py:356-430`), for files of one shape: the loss along a sampled
    label trajectory, the share of its steps whose top-scoring pair is admissible, the normalised RF distance of the argmax tree to the
    label tree and - with `with_likelihood` - the branch-optimised log-likelihoods of the label topology (`scores_ref`, `train.py:381`) and
    of the argmax tree (`scores`, `:387`) from the GPU likelihood (needs the real sequences in the files)."""
    from .rollout import reinforce_rollout
    batch = make_batch(files, rng)
    sup = supervise_rollout(batch, agent, env, eval=True, branch_optimize=with_likelihood)
    loss = balanced_elu_loss(sup, epoch=epoch, ratio_factor=ratio_factor)
    hits = []
    for t, lg in enumerate(sup[0]):
        top = lg.argmax(dim=1).cpu().tolist()
        hits.append([top[b] in sup[1][t][b] for b in range(len(top))])
    _, _, scores, _ = reinforce_rollout(batch, agent, env, cfgs, eval=True, argmax=True, branch_optimize=with_likelihood)
    rf = [normalized_rf(batch["label_newick"][b], env.states[b].subtrees[0].utree_op_str) for b in range(len(files))]
    out = {"loss": loss["loss"], "precision": loss["precision"], "argmax_in_action_set": float(np.mean(hits)),
           "normalized_rf": rf, "normalized_rf_mean": float(np.mean(rf))}
    if with_likelihood:
        ref, got = sup.scores.double().cpu(), scores.double().cpu()
        out.update({"llh_label_tree": ref.tolist(), "llh_argmax_tree": got.tolist(), "llh_diff_mean": float((ref - got).mean())})
    return out


def evaluate_dir(data_dir: str, agent, cfgs, device, batch: int = 32, limit: Optional[int] = None, epoch: int = 0,
                 ratio_factor: float = 0.5, with_likelihood: bool = False, seed: int = 0) -> dict:
    """`eval_on_dataset` over a directory tree of `<name>.phy` + `<name>.tre` pairs (the layout of the reference's
    data_gen/data/test/len*/taxa*): files are grouped by directory (one shape per directory, like `PhySampler`, `train.py:297-299`)
    and evaluated `batch` at a time; means are over files."""
    import os
    from .environment import PhyInferEnv
    rng = _random.Random(seed)
    groups = {}
    for root, _, names in sorted(os.walk(data_dir)):
        for n in sorted(names):
            if n.endswith(".phy") and os.path.exists(os.path.join(root, n[:-4] + ".tre")):
                groups.setdefault(root, []).append((os.path.join(root, n), os.path.join(root, n[:-4] + ".tre")))
    per_dir, rows = {}, []
    for root, files in groups.items():
        files = files[:limit] if limit else files
        acc = []
        for k in range(0, len(files), batch):
            chunk = files[k:k + batch]
            env = PhyInferEnv(cfgs, device)
            res = evaluate(chunk, agent, env, cfgs=cfgs, epoch=epoch, ratio_factor=ratio_factor, rng=rng, with_likelihood=with_likelihood)
            acc.append((len(chunk), res))
        tot = sum(n for n, _ in acc)
        row = {"files": tot, "loss": sum(n * r["loss"] for n, r in acc) / tot, "precision": sum(n * r["precision"] for n, r in acc) / tot,
               "argmax_in_action_set": sum(n * r["argmax_in_action_set"] for n, r in acc) / tot,
               "normalized_rf_mean": sum(n * r["normalized_rf_mean"] for n, r in acc) / tot}
        if with_likelihood:
            row["llh_diff_mean"] = sum(n * r["llh_diff_mean"] for n, r in acc) / tot
        per_dir[os.path.relpath(root, data_dir)] = row
        rows.append(row)
    if not rows:
        raise NnjError(f"{data_dir}: no <name>.phy / <name>.tre pairs found")
    n = sum(r["files"] for r in rows)
    return {"files": n, "loss": sum(r["files"] * r["loss"] for r in rows) / n,
            "normalized_rf_mean": sum(r["files"] * r["normalized_rf_mean"] for r in rows) / n, "by_directory": per_dir}


def main(argv=None):
    """Validation of a checkpoint on labelled alignments: `python -m neuralnj_b200.supervise --config_path cfg.yaml --data_dir DIR`."""
    import argparse
    import json
    from .config import empty_config, inference_config
    from .rollout import _load_policy
    ap = argparse.ArgumentParser(description="NeuralNJ validation metrics (train.py eval_on_dataset) on the B200 path")
    ap.add_argument("--config_path", type=str, default="")
    ap.add_argument("--data_dir", type=str, required=True)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--limit", type=int, default=None, help="files per directory")
    ap.add_argument("--epoch", type=int, default=0, help="position in the K-ratio schedule of the loss (train.py:495)")
    ap.add_argument("--likelihood", action="store_true", help="also score label and argmax trees with the GPU likelihood")
    ap.add_argument("--precision", type=str, default=None, choices=["fp32", "bf16x3", "bf16"])
    args = ap.parse_args(argv)
    if args.config_path:
        cfgs = empty_config()
        cfgs.merge_from_file(args.config_path)
    else:       # no YAML given: the shipped inference model (config/finetune_reinforce_search_example.yaml:24-30)
        cfgs = inference_config()
    device = torch.device("cuda:0")
    agent = _load_policy(cfgs, device, args.precision)
    ratio_factor = getattr(cfgs, "ratio_factor", 0.5)
    out = evaluate_dir(args.data_dir, agent, cfgs, device, batch=args.batch, limit=args.limit, epoch=args.epoch,
                       ratio_factor=ratio_factor, with_likelihood=args.likelihood)
    print(json.dumps(out, indent=1))
    return out


if __name__ == "__main__":
    main()
