"""`PhyInferEnv` — host-side tree bookkeeping with the reference's interface (environment.py:444-898).

What callers of the inference path rely on is kept: `init_states`, `step`, `evaluate_loglikelihood`,
`dump_end_trees`, `get_current_trees`, `action_to_indices`, and the attributes `states`,
`state_tensor`, `init_state_tensor`, `tree_pairs_dict`, `action_indices_dict`, `batch_seqs`,
`seq_keys`, `batch_size`.  Tree objects expose the fields the reference's formatting code reads
(`left_tree_data` / `right_tree_data` dicts with "tree" and "branch_length", `seq_indices`,
`log_score`, `utree_op_str`, ...).  The tensor half of `step` (gather, aggregate, re-index;
environment.py:760-835) runs in one CUDA call (`agent.merge_state`).  `replay_merges` rebuilds the
same host state from a device-produced merge list (fused rollout).

`branch_optimize=True` (`optimize_branch_length_sequential`, environment.py:648-671: RAxML-NG through the native raxmlpy
binding) scores the finished trees with the GPU likelihood of `neuralnj_b200.likelihood` (GTR+I+G4, csrc/nnj_llh.cu): all
trees of the batch in one call.  Out of scope: label-tree supervision (training).
"""
from __future__ import annotations

import functools
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from .treeutil import treestr_to_tuples

PLACEHOLDER_BRANCH = 0.12345      # environment.py:284
UNSCORED = -111111                # environment.py:686
EVOLUTION_MODEL = "GTR+I+G"       # environment.py:44

CHARACTERS_MAPS = {
    "DNA": dict(A=[1., 0., 0., 0.], C=[0., 1., 0., 0.], G=[0., 0., 1., 0.], T=[0., 0., 0., 1.], N=[1., 1., 1., 1.]),
    "DNA_WITH_GAP": {"A": [1., 0., 0., 0.], "C": [0., 1., 0., 0.], "G": [0., 0., 1., 0.], "T": [0., 0., 0., 1.],
                     "-": [1., 1., 1., 1.], "N": [1., 1., 1., 1.]},
}
CHARACTERS_MAPS["RNA"] = {("U" if k == "T" else k): v for k, v in CHARACTERS_MAPS["DNA"].items()}
CHARACTERS_MAPS["RNA_WITH_GAP"] = {("U" if k == "T" else k): v for k, v in CHARACTERS_MAPS["DNA_WITH_GAP"].items()}


class PhyloTree:
    """Rooted binary (sub)tree.  The child holding the smaller first leaf index is the left child (environment.py:79-83)."""

    def __init__(self, at_root, left_tree_data=None, right_tree_data=None, root_seq_data=None, name="", device="cpu"):
        if root_seq_data is None and (left_tree_data is None or right_tree_data is None):
            raise ValueError("PhyloTree needs either a leaf index or two children")
        self.device = device
        self.at_root = at_root
        self.name = name
        if root_seq_data is not None:
            self.left_tree_data = self.right_tree_data = None
            self.seq_indices = list(root_seq_data)
            self.total_mutations = 0
            self.log_score = 0
        else:
            if left_tree_data["tree"].seq_indices[0] > right_tree_data["tree"].seq_indices[0]:
                left_tree_data, right_tree_data = right_tree_data, left_tree_data
            self.left_tree_data, self.right_tree_data = left_tree_data, right_tree_data
            self.seq_indices = sorted(left_tree_data["tree"].seq_indices + right_tree_data["tree"].seq_indices)
            self.root_seq = None
            self.log_score = None
        self.seq_indices_str = str(self.seq_indices)
        self.root_rep_str = self.seq_indices_str
        self.min_seq_index = self.seq_indices[0]

    @property
    def is_leaf(self) -> bool:
        return self.left_tree_data is None

    @property
    def is_internal(self) -> bool:
        return self.left_tree_data is not None

    def children(self):
        return () if self.is_leaf else (self.left_tree_data, self.right_tree_data)

    def to_unrooted_tree(self) -> "UnrootedPhyloTree":
        """Edge-length table with the two root edges fused into one (environment.py:154-200)."""
        lengths = {}
        stack = [self]
        while stack:
            node = stack.pop()
            for child in node.children():
                a, b = child["tree"].seq_indices_str, node.seq_indices_str
                lengths[(a, b)] = lengths[(b, a)] = child["branch_length"]
                stack.append(child["tree"])
        top = self.seq_indices_str
        l, r = self.left_tree_data["tree"].seq_indices_str, self.right_tree_data["tree"].seq_indices_str
        ll, lr = lengths.pop((top, l)), lengths.pop((top, r))
        lengths.pop((l, top)); lengths.pop((r, top))
        lengths[(l, r)] = lengths[(r, l)] = (ll + lr) if (ll is not None and lr is not None) else None
        return UnrootedPhyloTree(self.log_score, self.left_tree_data, self.right_tree_data, lengths, self.seq_indices)


class UnrootedPhyloTree:
    def __init__(self, log_score, left_tree_data, right_tree_data, branch_length, seq_indices, name=""):
        self.left_tree_data, self.right_tree_data = left_tree_data, right_tree_data
        self.branch_length = branch_length
        self.log_score = log_score
        self.seq_indices = seq_indices
        self.topo_repr = format_rtree_topology(self, True, None)
        self.name = name


def format_rtree_topology(tree, at_root=False, sequence_keys=None) -> str:
    """Newick topology without lengths (environment.py:263-277)."""
    if tree.left_tree_data is None:
        i = tree.seq_indices[0]
        return f"{sequence_keys[i] if sequence_keys else i}"
    l = format_rtree_topology(tree.left_tree_data["tree"], False, sequence_keys)
    r = format_rtree_topology(tree.right_tree_data["tree"], False, sequence_keys)
    return f"({l}, {r});" if at_root else f"({l}, {r})"


def format_rtree(tree, at_root=False, branch_length=None, sequence_keys=None) -> str:
    """Newick with branch lengths; missing lengths print as 0.12345 (environment.py:280-304)."""
    if sequence_keys is None:
        raise ValueError("format_rtree needs sequence_keys")
    b = PLACEHOLDER_BRANCH if branch_length is None else branch_length
    if tree.left_tree_data is None:
        return f"{sequence_keys[tree.seq_indices[0]]}:{b}"
    l = format_rtree(tree.left_tree_data["tree"], False, tree.left_tree_data["branch_length"], sequence_keys)
    r = format_rtree(tree.right_tree_data["tree"], False, tree.right_tree_data["branch_length"], sequence_keys)
    return f"({l}, {r});" if at_root else f"({l}, {r}):{b}"


def assign_branch_length_rtree(tree, tree_tuple) -> None:
    """Write the lengths of a (child, len, child, len) tuple tree back onto the PhyloTree (environment.py:307-321)."""
    if tree.left_tree_data is None:
        return
    lt, lb, rt, rb = tree_tuple
    tree.left_tree_data["branch_length"], tree.right_tree_data["branch_length"] = lb, rb
    assign_branch_length_rtree(tree.left_tree_data["tree"], lt)
    assign_branch_length_rtree(tree.right_tree_data["tree"], rt)


class PhylogeneticTreeState:
    def __init__(self, subtrees: List):
        self.subtrees = subtrees
        self.num_trees = len(subtrees)
        self.is_done = self.num_trees == 1
        if isinstance(subtrees[0], PhyloTree):
            self.is_initial = all(t.left_tree_data is None for t in subtrees)
            self.last_state = False
            self.log_score = None
        else:
            self.is_initial = False
            self.last_state = True
            self.log_score = subtrees[0].log_score


@functools.lru_cache(maxsize=8)
def pair_tables(n_max: int):
    """tree_pairs_dict / action_indices_dict for n = 2..n_max (environment.py:455-462).  Read-only tables, built once per taxon
    count (5 ms at 50 taxa: a third of a single-alignment rollout if rebuilt by every init_states call)."""
    pairs, index = {}, {}
    for n in range(2, n_max + 1):
        lst = [(i, j) for i in range(n) for j in range(i + 1, n)]
        pairs[n] = lst
        index[n] = {p: k for k, p in enumerate(lst)}
    return pairs, index


class PhyInferEnv:
    def __init__(self, cfg, device):
        self.device = device
        self.chars_dict = CHARACTERS_MAPS[cfg.env.sequence_type]
        self.states = None
        self.state_tensor = None

    # ------------------------------------------------------------------ setup
    def init_states(self, batch_seqs, seq_keys, seq_arrays, label_trees=None, step_action=False):
        if label_trees is not None:
            raise NotImplementedError("label-tree supervision belongs to training (train.py), outside the inference hot path")
        self.batch_seqs, self.seq_keys = batch_seqs, seq_keys
        self.batch_size = len(batch_seqs)
        n = len(batch_seqs[0])
        self.tree_pairs_dict, self.action_indices_dict = pair_tables(n)
        self.states = [
            PhylogeneticTreeState([PhyloTree(False, root_seq_data=[i], device=self.device, name=seq_keys[b][i]) for i in range(len(batch_seqs[b]))])
            for b in range(self.batch_size)
        ]
        self.init_state_tensor = seq_arrays
        self.state_tensor = None
        self.label_trees, self.mom_maps, self.batch_action_set_step = None, dict(), []

    def action_to_indices(self, actions):
        table = self.tree_pairs_dict[self.states[0].num_trees]
        return torch.tensor([table[int(a)] for a in actions], dtype=torch.long)

    # ------------------------------------------------------------------ host half of a step
    def _join(self, b: int, i: int, j: int, edge=(None, None)) -> Tuple[PhyloTree, bool]:
        st = self.states[b]
        if st.is_done:
            raise RuntimeError("step() on a finished tree")
        last = len(st.subtrees) == 2
        tree = PhyloTree(last, {"tree": st.subtrees[i], "branch_length": edge[0]}, {"tree": st.subtrees[j], "branch_length": edge[1]}, name="")
        return tree, last

    def _finish_unscored(self, b: int, tree: PhyloTree) -> None:
        """`optimize_branch_length_no_br` (environment.py:674-686): placeholder lengths, score -111111."""
        rtree_str = format_rtree(tree, True, None, self.seq_keys[b])
        tup = treestr_to_tuples(rtree_str)
        assign_branch_length_rtree(tree, tup)
        tree.rtree_op_tuple = tree.utree_op_tuple = tup
        tree.utree_op_str = rtree_str
        tree.log_score = UNSCORED

    def _finish_scored(self, trees: List[PhyloTree]) -> None:
        """`optimize_branch_length_sequential` (environment.py:648-671) for the whole batch at once: branch lengths and the
        GTR+I+G model parameters are optimised on the GPU (likelihood.TreeLikelihood), the optimised lengths go back onto the
        rooted trees and `log_score` is the maximised log-likelihood."""
        from . import likelihood as LH
        R = len(self.batch_seqs[0])
        masks_all = LH.onehot_to_masks(self.init_state_tensor) if self.init_state_tensor is not None else None
        pats, wts, children = [], [], []
        for b, tree in enumerate(trees):
            masks = masks_all[b] if masks_all is not None else LH.sequences_to_masks(self.batch_seqs[b])
            p, w = LH.compress_patterns(masks)
            pats.append(p); wts.append(w)
            ch = []

            def visit(t):
                if t.is_leaf:
                    return t.seq_indices[0]
                a, c = visit(t.left_tree_data["tree"]), visit(t.right_tree_data["tree"])
                ch.append((a, c))
                return R + len(ch) - 1
            visit(tree)
            children.append(ch)
        Lp = max(p.shape[1] for p in pats)
        P = np.full((len(trees), R, Lp), 15, dtype=np.uint8)     # padding patterns: fully undetermined, weight 0
        W = np.zeros((len(trees), Lp))
        for b, (p, w) in enumerate(zip(pats, wts)):
            P[b, :, :p.shape[1]] = p
            W[b, :len(w)] = w
        children = np.asarray(children, dtype=np.int32)
        eng = LH.TreeLikelihood(P, W, device=self.device)
        freqs = np.stack([LH.empirical_freqs(masks_all[b] if masks_all is not None else LH.sequences_to_masks(self.batch_seqs[b])) for b in range(len(trees))])
        model = LH.SubstModel(EVOLUTION_MODEL, freqs, len(trees))
        t, _, ll = eng.optimize_all(children, np.full((len(trees), 2 * R - 2), LH.BRLEN_DEFAULT), model)
        for b, tree in enumerate(trees):
            keys = self.seq_keys[b]
            tree.rtree_op_tuple = LH.tuples_with_lengths(children[b], t[b], keys, unrooted=False)
            tree.utree_op_tuple = LH.tuples_with_lengths(children[b], t[b], keys, unrooted=True)
            tree.utree_op_str = LH.tuples_to_newick(tree.utree_op_tuple)
            assign_branch_length_rtree(tree, tree.rtree_op_tuple)
            tree.log_score = float(ll[b])

    def _advance_host(self, ij: Sequence[Tuple[int, int]], edge_actions, branch_optimize: bool) -> bool:
        done = False
        new_states = []
        joined = [self._join(b, i, j, edge_actions[b] if edge_actions is not None else (None, None)) for b, (i, j) in enumerate(ij)]
        if joined and joined[0][1] and branch_optimize:
            self._finish_scored([t for t, _ in joined])
        for b, (i, j) in enumerate(ij):
            tree, last = joined[b]
            if last:
                if not branch_optimize:
                    self._finish_unscored(b, tree)
                ut = tree.to_unrooted_tree()
                ut.rtree_op_tuple, ut.utree_op_tuple, ut.utree_op_str = tree.rtree_op_tuple, tree.utree_op_tuple, tree.utree_op_str
                ns = PhylogeneticTreeState([ut])
            else:
                subs = self.states[b].subtrees
                subs[i] = tree          # slot i <- merged subtree, slot j removed (environment.py:735-738)
                subs.pop(j)
                ns = PhylogeneticTreeState(subs)
            new_states.append(ns)
            done = ns.is_done
        self.states = new_states
        self.batch_action_set_step = [] if done else [None] * len(ij)
        return done

    # ------------------------------------------------------------------ reference API
    def step(self, actions, edge_actions, parallel=True, branch_optimize=True, agent=None, step_action=False):
        n = self.states[0].num_trees
        table = self.tree_pairs_dict[n]
        ij = [tuple(int(v) for v in table[int(a)]) for a in actions]
        done = self._advance_host(ij, edge_actions, branch_optimize)
        if not done:
            ij_t = torch.tensor(ij, dtype=torch.int32, device=self.state_tensor.device)
            if agent is not None and hasattr(agent, "merge_state"):
                self.state_tensor = agent.merge_state(self.state_tensor, ij_t)
            else:
                self.state_tensor = self._merge_generic(ij_t.long(), agent)
        return done

    def _merge_generic(self, ij: torch.Tensor, agent) -> torch.Tensor:
        """Reference formulation with torch ops for foreign agents: mean (agent=None) or agent.aggregate."""
        st = self.state_tensor
        B, n = st.shape[:2]
        ar = torch.arange(B, device=st.device)
        xi, xj = st[ar, ij[:, 0]].unsqueeze(1), st[ar, ij[:, 1]].unsqueeze(1)
        new = (xi + xj) / 2 if agent is None else agent.aggregate(xi, xj, (ij[:, 0], ij[:, 1]), batchwise_ij_indices=True)
        rows = []
        for b in range(B):
            i, j = int(ij[b, 0]), int(ij[b, 1])
            keep = [r for r in range(n) if r != j]
            xb = st[b, keep].clone()
            xb[keep.index(i)] = new[b, 0]
            rows.append(xb)
        return torch.stack(rows)

    def replay_merges(self, merges, branch_optimize: bool = False) -> None:
        """Apply a device-produced merge list [B,R-1,2] to the host trees (fused rollout epilogue); with `branch_optimize`
        the finished trees are scored like `step(..., branch_optimize=True)` does."""
        merges = np.asarray(merges.cpu() if torch.is_tensor(merges) else merges)
        for t in range(merges.shape[1]):
            self._advance_host([(int(a), int(b)) for a, b in merges[:, t]], None, branch_optimize and t == merges.shape[1] - 1)
        self.state_tensor = None

    def dump_end_trees(self):
        trees = [s.subtrees[0] for s in self.states]
        return trees, [t.log_score for t in trees]

    def evaluate_loglikelihood(self, get_all_tree=False):
        for s in self.states:
            if not s.is_done:
                raise RuntimeError("evaluate_loglikelihood() before the rollout finished")
        scores = [s.log_score for s in self.states]
        score_t = torch.from_numpy(np.array(scores)).to(self.device)
        if get_all_tree:
            ends = [s.subtrees[0] for s in self.states]
            return score_t, [t.rtree_op_tuple for t in ends], [t.utree_op_tuple for t in ends], [t.utree_op_str for t in ends]
        best = self.states[scores.index(max(scores))].subtrees[0]
        return score_t, best.rtree_op_tuple, best.utree_op_tuple, best.utree_op_str

    def get_current_trees(self):
        return [[format_rtree_topology(t, at_root=True, sequence_keys=None) for t in s.subtrees] for s in self.states]

    def _seq2array(self, seq):
        return np.array([self.chars_dict[x] for x in seq])
