"""CPU oracle for the NeuralNJ inference hot path.  TEST INFRASTRUCTURE ONLY.

This file is a functional restatement (torch CPU, fp32 by default, fp64 on
request) of the reference algorithm for the path named in BASELINE.json:
MSA axial encoder + learned neighbour-joining loop with the stale-score cache.
It is the checker for tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under neuralnj_b200/ may import
it: the product path is CUDA only.

Parity status: PINNED.  oracle/make_golden.py runs the unmodified reference
(imported from /root/reference with stub modules for its absent third-party
imports) and this restatement on the same MSAs and seeds; the reference's
outputs are committed under tests/golden/ and tests/test_oracle.py checks this
file against them (merge lists identical, logits to 1e-5 of the step's
max|logit|, Newick identical).  The reference ships no tests or golden vectors
for this path, so the executed reference is the pin.

Reference citations (relative to /root/reference):
  encode            model.py:67-88, msa_modules.py:62-125,128-151
  row attention     axial_attention.py:31-138   (chunked form :35-64)
  column attention  axial_attention.py:166-255
  aggregate         model.py:102-155
  pair scores       model.py:90-99, 158-209
  cache index map   utils.py:213-251
  select            finetune_rl_search.py:140-160
  merge / re-index  environment.py:760-835
  Newick            environment.py:280-304, 79-83
"""
from __future__ import annotations

import math

import numpy as np
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

MAX_TOKENS_PER_MSA = 1024  # model.py:34 (hard-coded 7th ctor argument)
NEG_FILL = -10000.0        # axial_attention.py:102, :223
LN_EPS = 1e-5              # torch.nn.LayerNorm default, msa_modules.py:107

State = Dict[str, torch.Tensor]


# --------------------------------------------------------------------------
# weights
# --------------------------------------------------------------------------
def state_dict_keys(num_layers: int = 6) -> List[str]:
    """The reference state_dict keys (172 for 6 layers) in construction order (model.py:12-60)."""
    keys = []
    for l in range(num_layers):
        for blk in ("row_self_attention", "column_self_attention"):
            for p in ("k_proj", "v_proj", "q_proj", "out_proj"):
                keys += [f"seq_emb_layers.{l}.{blk}.layer.{p}.weight",
                         f"seq_emb_layers.{l}.{blk}.layer.{p}.bias"]
            keys += [f"seq_emb_layers.{l}.{blk}.layer_norm.weight",
                     f"seq_emb_layers.{l}.{blk}.layer_norm.bias"]
        for p in ("fc1", "fc2"):
            keys += [f"seq_emb_layers.{l}.feed_forward_layer.layer.{p}.weight",
                     f"seq_emb_layers.{l}.feed_forward_layer.layer.{p}.bias"]
        keys += [f"seq_emb_layers.{l}.feed_forward_layer.layer_norm.weight",
                 f"seq_emb_layers.{l}.feed_forward_layer.layer_norm.bias"]
    for p in ("embed.0", "embed.2", "h_linear_last", "g_linear_last",
              "g_attn_q", "g_attn_k", "s_out.0", "s_out.2"):
        keys += [f"{p}.weight", f"{p}.bias"]
    return keys


def _linear_init(out_f: int, in_f: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """torch.nn.Linear.reset_parameters, consuming the global RNG identically."""
    lin = torch.nn.Linear(in_f, out_f)
    return lin.weight.detach().clone(), lin.bias.detach().clone()


def init_state_dict(seed: int = 0, embed_dim: int = 64, num_layers: int = 6,
                    vocab: int = 4, patch: int = 1) -> State:
    """Default-initialised weights equal to `torch.manual_seed(seed); PhyloATTN(cfgs)`.

    The RNG is consumed in the reference's module construction order
    (model.py:25-59; axial_attention.py:24-28,161-165; msa_modules.py:142-143):
    per layer row{k,v,q,out}, column{k,v,q,out}, fc1, fc2; then embed.0,
    embed.2, h_linear_last, g_linear_last, g_attn_q, g_attn_k, s_out.0, s_out.2.
    LayerNorm consumes no randomness.
    """
    torch.manual_seed(seed)
    D, Fh = embed_dim, 4 * embed_dim
    sd: State = {}
    for l in range(num_layers):
        for blk in ("row_self_attention", "column_self_attention"):
            for p in ("k_proj", "v_proj", "q_proj", "out_proj"):
                w, b = _linear_init(D, D)
                sd[f"seq_emb_layers.{l}.{blk}.layer.{p}.weight"] = w
                sd[f"seq_emb_layers.{l}.{blk}.layer.{p}.bias"] = b
        w, b = _linear_init(Fh, D)
        sd[f"seq_emb_layers.{l}.feed_forward_layer.layer.fc1.weight"] = w
        sd[f"seq_emb_layers.{l}.feed_forward_layer.layer.fc1.bias"] = b
        w, b = _linear_init(D, Fh)
        sd[f"seq_emb_layers.{l}.feed_forward_layer.layer.fc2.weight"] = w
        sd[f"seq_emb_layers.{l}.feed_forward_layer.layer.fc2.bias"] = b
        for blk in ("row_self_attention", "column_self_attention", "feed_forward_layer"):
            sd[f"seq_emb_layers.{l}.{blk}.layer_norm.weight"] = torch.ones(D)
            sd[f"seq_emb_layers.{l}.{blk}.layer_norm.bias"] = torch.zeros(D)
    for name, (o, i) in (("embed.0", (D, vocab * patch)), ("embed.2", (D, D)),
                         ("h_linear_last", (D, D)), ("g_linear_last", (D, D)),
                         ("g_attn_q", (D, D)), ("g_attn_k", (D, D)),
                         ("s_out.0", (D, D)), ("s_out.2", (1, D))):
        w, b = _linear_init(o, i)
        sd[f"{name}.weight"] = w
        sd[f"{name}.bias"] = b
    return {k: sd[k] for k in state_dict_keys(num_layers)}


def _lin(sd: State, name: str, x: torch.Tensor) -> torch.Tensor:
    return F.linear(x, sd[name + ".weight"], sd[name + ".bias"])


def cast_state_dict(sd: State, dtype: torch.dtype) -> State:
    return {k: v.to(dtype) for k, v in sd.items()}


# --------------------------------------------------------------------------
# encoder
# --------------------------------------------------------------------------
def _row_logits(sd: State, pre: str, x: torch.Tensor, scaling: float,
                pad: Optional[torch.Tensor], H: int) -> torch.Tensor:
    """Tied row-attention logits of a row chunk.  axial_attention.py:66-105.

    x [r,C,B,D]; pad bool [B,r,C] or None.  Returns [H,B,C,C].
    """
    r, C, B, D = x.shape
    dh = D // H
    q = _lin(sd, pre + ".q_proj", x).view(r, C, B, H, dh)
    k = _lin(sd, pre + ".k_proj", x).view(r, C, B, H, dh)
    q = q * scaling
    if pad is not None:
        q = q * (1 - pad.permute(1, 2, 0).unsqueeze(3).unsqueeze(4).to(q))
    w = torch.einsum("rinhd,rjnhd->hnij", q, k)
    if pad is not None:
        w = w.masked_fill(pad[:, 0].unsqueeze(0).unsqueeze(2), NEG_FILL)
    return w


def _row_update(sd: State, pre: str, x: torch.Tensor, probs: torch.Tensor, H: int) -> torch.Tensor:
    """axial_attention.py:107-117."""
    r, C, B, D = x.shape
    v = _lin(sd, pre + ".v_proj", x).view(r, C, B, H, D // H)
    ctx = torch.einsum("hnij,rjnhd->rinhd", probs, v).contiguous().view(r, C, B, D)
    return _lin(sd, pre + ".out_proj", ctx)


def row_attention(sd: State, pre: str, x: torch.Tensor, pad: Optional[torch.Tensor],
                  H: int, chunked: bool = True) -> torch.Tensor:
    """Tied row self-attention over x [R,C,B,D].  axial_attention.py:119-138.

    `chunked=True` follows the inference (no_grad) path :35-64: rows in chunks
    of max(1, 1024 // C), logits summed over chunks before the softmax.
    """
    R, C, B, D = x.shape
    scaling = (D // H) ** -0.5 / math.sqrt(R)  # align_scaling :31-33, full R also when chunked :44
    if chunked and R * C > MAX_TOKENS_PER_MSA:
        step = max(1, MAX_TOKENS_PER_MSA // C)
        logits = 0
        for s in range(0, R, step):
            logits = logits + _row_logits(sd, pre, x[s:s + step], scaling,
                                          pad[:, s:s + step] if pad is not None else None, H)
        probs = logits.softmax(-1)
        return torch.cat([_row_update(sd, pre, x[s:s + step], probs, H)
                          for s in range(0, R, step)], 0)
    probs = _row_logits(sd, pre, x, scaling, pad, H).softmax(-1)
    return _row_update(sd, pre, x, probs, H)


def _col_block(sd: State, pre: str, x: torch.Tensor, pad: Optional[torch.Tensor], H: int) -> torch.Tensor:
    """Column attention of a column chunk.  axial_attention.py:190-237."""
    R, C, B, D = x.shape
    dh = D // H
    if R == 1:
        return _lin(sd, pre + ".out_proj", _lin(sd, pre + ".v_proj", x))
    q = _lin(sd, pre + ".q_proj", x).view(R, C, B, H, dh) * (dh ** -0.5)
    k = _lin(sd, pre + ".k_proj", x).view(R, C, B, H, dh)
    v = _lin(sd, pre + ".v_proj", x).view(R, C, B, H, dh)
    w = torch.einsum("icnhd,jcnhd->hcnij", q, k)
    if pad is not None:
        w = w.masked_fill(pad.permute(2, 0, 1).unsqueeze(0).unsqueeze(3), NEG_FILL)
    p = w.softmax(-1)
    ctx = torch.einsum("hcnij,jcnhd->icnhd", p, v).contiguous().view(R, C, B, D)
    return _lin(sd, pre + ".out_proj", ctx)


def column_attention(sd: State, pre: str, x: torch.Tensor, pad: Optional[torch.Tensor],
                     H: int, chunked: bool = True) -> torch.Tensor:
    """axial_attention.py:239-255 (chunks of max(1, 1024 // R) columns :166-188)."""
    R, C, B, D = x.shape
    if chunked and R * C > MAX_TOKENS_PER_MSA:
        step = max(1, MAX_TOKENS_PER_MSA // R)
        return torch.cat([_col_block(sd, pre, x[:, s:s + step],
                                     pad[:, :, s:s + step] if pad is not None else None, H)
                          for s in range(0, C, step)], 1)
    return _col_block(sd, pre, x, pad, H)


def encode(sd: State, data: torch.Tensor, seq_mask: torch.Tensor, num_heads: int = 8,
           patch: int = 1, chunked: bool = True, return_layers: bool = False):
    """`PhyloATTN.encode_zxr` (model.py:67-88).

    data int8 [B,R,L,4]; seq_mask bool [B,L] (True = padded column).
    Returns fp [B,R,C,D]  (and the per-layer outputs when return_layers).
    """
    dt = sd["embed.0.weight"].dtype
    x = data.to(dt)
    B, R, L, E = x.shape
    C = math.ceil(L / patch)
    x = x.reshape(B, R, C, patch * E)
    x = _lin(sd, "embed.2", F.gelu(_lin(sd, "embed.0", x)))
    pad = seq_mask[:, ::patch].unsqueeze(1).expand(B, R, C)  # einops.repeat 'b s -> b n s'
    x = x.permute(1, 2, 0, 3)  # [R,C,B,D]
    layers = []
    num_layers = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("seq_emb_layers."))
    D = x.shape[-1]
    for l in range(num_layers):
        base = f"seq_emb_layers.{l}"
        for blk in ("row_self_attention", "column_self_attention", "feed_forward_layer"):
            y = F.layer_norm(x, (D,), sd[f"{base}.{blk}.layer_norm.weight"],
                             sd[f"{base}.{blk}.layer_norm.bias"], LN_EPS)
            pre = f"{base}.{blk}.layer"
            if blk == "row_self_attention":
                y = row_attention(sd, pre, y, pad, num_heads, chunked)
            elif blk == "column_self_attention":
                y = column_attention(sd, pre, y, pad, num_heads, chunked)
            else:
                y = _lin(sd, pre + ".fc2", F.gelu(_lin(sd, pre + ".fc1", y)))
            x = x + y  # dropout(p=0.4) is the identity in eval (msa_modules.py:119-120)
        if return_layers:
            layers.append(x.permute(2, 0, 1, 3).contiguous())
    out = x.permute(2, 0, 1, 3)
    return (out, layers) if return_layers else out


# --------------------------------------------------------------------------
# pair scorer / merger
# --------------------------------------------------------------------------
def aggregate(sd: State, nodes: torch.Tensor, x_i: torch.Tensor, x_j: torch.Tensor,
              i_idx: torch.Tensor, j_idx: torch.Tensor, patch_num: int) -> torch.Tensor:
    """`PhyloATTN.aggregate` (model.py:102-155).

    nodes [B,R',C,D] (the reference's implicit self.batch_input); x_i,x_j
    [B,N,C,D]; i_idx,j_idx int64 [B,N] = node indices masked out of the global
    attention for pair n of tree b (covers the reference's three index forms).
    """
    z = torch.sigmoid(_lin(sd, "h_linear_last", x_i - x_j))
    x = z * x_i + (1 - z) * x_j
    if nodes.size(1) > 2:
        D = nodes.size(-1)
        q = _lin(sd, "g_attn_q", x)
        k = _lin(sd, "g_attn_k", nodes)
        alpha = torch.einsum("bncd,brcd->bnr", q, k) / math.sqrt(D * patch_num)
        ninf = torch.full_like(alpha, 0.0)
        ninf.scatter_(2, i_idx.unsqueeze(-1), float("-inf"))
        ninf.scatter_(2, j_idx.unsqueeze(-1), float("-inf"))
        alpha = torch.softmax(alpha + ninf, dim=-1)
        xg = torch.einsum("bnr,brcd->bncd", alpha, nodes)
        w = torch.sigmoid(_lin(sd, "g_linear_last", xg))
        x = (1 - w) * x + w * xg
    return x


def pair_scores(sd: State, nodes: torch.Tensor, valid: torch.Tensor, i_idx: torch.Tensor,
                j_idx: torch.Tensor, patch_num: int, pair_chunk: int = 0) -> torch.Tensor:
    """`decode_gg` (model.py:90-99): sum over valid sites of s_out(aggregate(.)).

    valid [B,1,C] (1 = real column).  pair_chunk>0 scores the pairs in chunks
    (same per-pair arithmetic; bounds the [B,N,C,D] temporaries for large MSAs).
    """
    B, N = i_idx.shape
    step = pair_chunk if pair_chunk > 0 else N
    out = []
    ar = torch.arange(B, device=nodes.device).unsqueeze(1)
    for s in range(0, N, step):
        ii, jj = i_idx[:, s:s + step], j_idx[:, s:s + step]
        x = aggregate(sd, nodes, nodes[ar, ii], nodes[ar, jj], ii, jj, patch_num)
        sc = _lin(sd, "s_out.2", F.gelu(_lin(sd, "s_out.0", x))).squeeze(-1)
        out.append((sc * valid).sum(-1))
    return torch.cat(out, 1)


def pair_index(i: int, j: int, n: int) -> int:
    """Position of (i,j), i<j, in itertools.combinations(range(n),2) (environment.py:458)."""
    return i * n - i * (i + 1) // 2 + (j - i - 1)


def pair_list(n: int) -> List[Tuple[int, int]]:
    return [(i, j) for i in range(n) for j in range(i + 1, n)]


def score_indices_to_prev(a: int, b: int, n_new: int) -> List[int]:
    """Closed form of utils.get_score_indices_to_prev (utils.py:213-251) for one tree.

    (a,b), a<b, is the pair merged from the (n_new+1)-node list; entry p of the
    result says where pair p of the n_new-node list is found in
    cat([logits_prev (P_old), new_scores (n_new)]).
    """
    n_old = n_new + 1
    p_old = n_old * (n_old - 1) // 2
    out = []
    for ii in range(n_new):
        for jj in range(ii + 1, n_new):
            if ii == a:
                out.append(p_old + jj)
            elif jj == a:
                out.append(p_old + ii)
            else:
                out.append(pair_index(ii + (ii >= b), jj + (jj >= b), n_old))
    return out


# --------------------------------------------------------------------------
# Newick / trees (host side)
# --------------------------------------------------------------------------
class _Node:
    __slots__ = ("left", "right", "leaves", "name")

    def __init__(self, left=None, right=None, leaf: int = -1):
        if left is None:
            self.left = self.right = None
            self.leaves = [leaf]
        else:
            # PhyloTree.__init__: the child holding the smaller first leaf goes left (environment.py:79-83)
            if left.leaves[0] > right.leaves[0]:
                left, right = right, left
            self.left, self.right = left, right
            self.leaves = sorted(left.leaves + right.leaves)


def _fmt(node: _Node, keys: Sequence[str], root: bool) -> str:
    """format_rtree with the 0.12345 placeholder length (environment.py:280-304)."""
    if node.left is None:
        return f"{keys[node.leaves[0]]}:0.12345"
    l, r = _fmt(node.left, keys, False), _fmt(node.right, keys, False)
    return f"({l}, {r});" if root else f"({l}, {r}):0.12345"


def newick_from_merges(merges: Sequence[Tuple[int, int]], keys: Sequence[str]) -> str:
    """Replay a merge list on the host exactly as PhyInferEnv.step does (environment.py:735-738)."""
    sub = [_Node(leaf=i) for i in range(len(keys))]
    for (i, j) in merges:
        sub[i] = _Node(sub[i], sub[j])
        sub.pop(j)
    assert len(sub) == 1
    return _fmt(sub[0], keys, True)


def bipartitions(newick: str) -> set:
    """Non-trivial leaf bipartitions of an (un)rooted Newick string (ete3-free RF support)."""
    s = newick.strip().rstrip(";")
    pos = 0
    splits = []

    def parse():
        nonlocal pos
        if s[pos] == "(":
            pos += 1
            leaves = set()
            while True:
                leaves |= parse()
                while s[pos] == " ":
                    pos += 1
                if s[pos] == ",":
                    pos += 1
                    while s[pos] == " ":
                        pos += 1
                    continue
                assert s[pos] == ")"
                pos += 1
                break
            while pos < len(s) and s[pos] not in ",)":
                pos += 1  # skip label / :length
            splits.append(frozenset(leaves))
            return leaves
        st = pos
        while pos < len(s) and s[pos] not in ",):":
            pos += 1
        name = s[st:pos].strip()
        while pos < len(s) and s[pos] not in ",)":
            pos += 1
        return {name}

    allv = frozenset(parse())
    out = set()
    for sp in splits:
        if 1 < len(sp) < len(allv) - 1:
            other = allv - sp
            out.add(min(sp, other, key=lambda t: (len(t), sorted(t))))
    return out


def rf_distance(nw_a: str, nw_b: str) -> int:
    """Unrooted Robinson-Foulds distance = |symmetric difference of bipartitions|."""
    return len(bipartitions(nw_a) ^ bipartitions(nw_b))


# --------------------------------------------------------------------------
# rollout
# --------------------------------------------------------------------------
def rollout(sd: State, data: torch.Tensor, seq_mask: torch.Tensor, num_heads: int = 8,
            patch: int = 1, gumbel: Optional[torch.Tensor] = None, chunked: bool = True,
            pair_chunk: int = 0, state: Optional[torch.Tensor] = None,
            forced_merges: Optional[torch.Tensor] = None) -> dict:
    """One full tree build per batch entry: `reinforce_rollout(eval=True, ...)`.

    finetune_rl_search.py:107-175 + model.py:158-209 + environment.py:760-835.
    Argmax when `gumbel` is None; otherwise action = argmax(logits + gumbel[:,t,:P])
    (Gumbel-max replay of Categorical sampling).
    `state` may supply a pre-computed encoder output; `forced_merges` [B,R-1,2]
    teacher-forces the actions (used to compare logits along one trajectory).
    Returns merges int64 [B,R-1,2], logits (list of [B,P'] per step), actions,
    selected_log_ps [B,R-2], log_ps, state0.
    """
    with torch.no_grad():
        B, R, L, _ = data.shape
        C = math.ceil(L / patch)
        X = encode(sd, data, seq_mask, num_heads, patch, chunked) if state is None else state
        state0 = X
        dev = X.device                        # CPU for the oracle proper; "cuda" turns this restatement into the eager-GPU bar of bench.py
        valid = (~seq_mask[:, None, ::patch]).to(X.dtype)
        merges = torch.zeros(B, R - 1, 2, dtype=torch.int64)
        logits_all, logp_all, sel, acts = [], [], [], []
        prev_ij = None
        logits_prev = None
        ar = torch.arange(B, device=dev)
        for t in range(R - 1):
            n = X.size(1)
            if logits_prev is None:
                ii, jj = torch.triu_indices(n, n, offset=1, device=dev)
                logits = pair_scores(sd, X, valid, ii.unsqueeze(0).expand(B, -1),
                                     jj.unsqueeze(0).expand(B, -1), C, pair_chunk)
            else:
                r_all = torch.arange(n, device=dev).unsqueeze(0).expand(B, n)
                a_i = prev_ij[:, :1].expand(B, n)
                lo, hi = torch.minimum(a_i, r_all), torch.maximum(a_i, r_all)
                new = pair_scores(sd, X, valid, lo, hi, C)  # includes the unused self pair (model.py:186-197)
                idx = torch.tensor([score_indices_to_prev(int(a), int(b), n) for a, b in prev_ij.tolist()],
                                   dtype=torch.int64, device=dev)
                logits = torch.gather(torch.cat([logits_prev, new], -1), 1, idx)
            logp = torch.log_softmax(logits, -1)
            if forced_merges is not None:
                fm = forced_merges[:, t]
                a = torch.tensor([pair_index(int(i), int(j), n) for i, j in fm.tolist()], device=dev)
            elif gumbel is None:
                a = torch.argmax(logits, -1)
            else:
                a = torch.argmax(logits + gumbel[:, t, :logits.size(1)].to(logits.dtype), -1)
            pl = pair_list(n)
            ij = torch.tensor([pl[k] for k in a.tolist()], dtype=torch.int64, device=dev)
            merges[:, t] = ij.cpu()
            logits_all.append(logits)
            acts.append(a)
            if n == 2:
                break
            logp_all.append(logp)
            sel.append(logp.gather(1, a.unsqueeze(1)))
            # merge: new node from the pre-merge node set; slot i <- new, slot j removed
            new_node = aggregate(sd, X, X[ar, ij[:, 0]].unsqueeze(1), X[ar, ij[:, 1]].unsqueeze(1),
                                 ij[:, :1], ij[:, 1:], C)
            rows = []
            for b in range(B):
                i, j = ij[b].tolist()
                xb = X[b].clone()
                xb[i] = new_node[b, 0]
                rows.append(torch.cat([xb[:j], xb[j + 1:]], 0))
            X = torch.stack(rows, 0)
            prev_ij = ij
            logits_prev = logits
        return {
            "merges": merges, "logits": logits_all, "actions": acts, "log_ps": logp_all,
            "selected_log_ps": torch.cat(sel, 1) if sel else torch.zeros(B, 0, device=dev),
            "state0": state0,
        }


# --------------------------------------------------------------------------
# inputs
# --------------------------------------------------------------------------
_ONEHOT = {"A": (1, 0, 0, 0), "C": (0, 1, 0, 0), "G": (0, 0, 1, 0), "T": (0, 0, 0, 1),
           "-": (1, 1, 1, 1), "N": (1, 1, 1, 1), "*": (0, 0, 0, 0)}  # phydata.py:38-46


def ranking_loss(logits_steps: Sequence[np.ndarray], set_index: np.ndarray, epoch: int = 0, ratio_factor: float = 0.5,
                 margin: float = 0.5) -> Tuple[float, float, List[float]]:
    """Balanced-ELU margin ranking loss of the pre-training step, restated (train.py:448-545, BALANCED_ELU_LOSS branch).

    `logits_steps[t]` [B, P_t] are the logits of the steps supervise_rollout records (all but the last, train.py:131-136);
    `set_index` int [B, T, width] holds each step's action set as pair indices, -1 padded (train.py:30-39).
    Per step: the complement of the action set is padded to the batch's widest one, its K = min(W, max(int(W * ratio), 8))
    largest scores are kept (padding masked out again), and elu(-(s - margin - u)) is averaged over all (set, kept) pairs of the
    batch; the loss is the mean over steps, `precision` the share of pairs with s > u over all steps.  float64 throughout.
    Returns (loss, precision, per-step losses)."""
    ratio = max(1 - 3 / (4 * 20) * epoch / 2, 1 / 4) * ratio_factor
    step_losses, right, pairs_total = [], 0.0, 0.0
    for t, lg in enumerate(logits_steps):
        lg = np.asarray(lg, dtype=np.float64)
        B, P = lg.shape
        members = [np.array([p for p in set_index[b, t] if p >= 0], dtype=np.int64) for b in range(B)]
        others = [np.setdiff1d(np.arange(P), m) for m in members]
        W = max(len(o) for o in others)
        K = min(W, max(int(W * ratio), 8))
        total, pairs = 0.0, 0.0
        for b in range(B):
            s = lg[b, members[b]]
            u = np.sort(lg[b, others[b]])[::-1][:K]
            x = -(s[:, None] - margin - u[None, :])
            total += float(np.where(x > 0, x, np.expm1(np.minimum(x, 0))).sum())
            pairs += s.size * u.size
            right += float((s[:, None] > u[None, :]).sum())
        step_losses.append(total / pairs)
        pairs_total += pairs
    return float(np.mean(step_losses)), right / pairs_total, step_losses


def load_phy(path: str):
    """PHYLIP (sequential or interleaved) -> (data int8 [1,R,L,4], mask bool [1,L], keys, seqs).

    phydata.py:499-548 (parse), :1249-1262 (taxa ordered by the integer suffix
    of their names), :98-123 (no padding is added when all rows have equal
    length, so the mask is all False).
    """
    import re
    with open(path) as f:
        n_taxa, n_sites = map(int, f.readline().split())
        names, seqs = [], {}
        block = []
        lines = [ln.strip() for ln in f]
    k = 0
    while k < len(lines) and lines[k]:
        parts = lines[k].split(maxsplit=1)
        if len(parts) > 1:
            names.append(parts[0])
            seqs[parts[0]] = parts[1].replace(" ", "").upper()
        k += 1
    idx = 0
    for ln in lines[k + 1:]:
        if ln:
            seqs[names[idx]] += ln.replace(" ", "").upper()
            idx += 1
        else:
            idx = 0
    assert len(names) == n_taxa
    pre = re.match(r"^([a-zA-Z]+)([0-9]+)$", names[0]).group(1)
    order = sorted(names, key=lambda s: int(s[len(pre):]))
    rows = []
    for nm in order:
        s = "".join(c if c in _ONEHOT else "-" for c in seqs[nm])
        assert len(s) == n_sites
        rows.append([_ONEHOT[c] for c in s])
    data = torch.tensor(rows, dtype=torch.int8).unsqueeze(0)
    mask = torch.zeros(1, n_sites, dtype=torch.bool)
    return data, mask, order, [seqs[nm] for nm in order]


def synthetic_msa(batch: int, taxa: int, sites: int, seed: int = 1234) -> torch.Tensor:
    """Config-2 generator: iid tokens over (A,C,G,T,gap)."""
    g = torch.Generator().manual_seed(seed)
    p = torch.tensor([0.28, 0.12, 0.12, 0.21, 0.27])
    tok = torch.multinomial(p, batch * taxa * sites, replacement=True, generator=g).view(batch, taxa, sites)
    table = torch.tensor([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1], [1, 1, 1, 1]], dtype=torch.int8)
    return table[tok]


def evolved_msa(batch: int, taxa: int, sites: int, seed: int = 7, rate: float = 0.08,
                gap: float = 0.05) -> torch.Tensor:
    """Seeded JC69-style simulation down a random binary tree (MSAs with phylogenetic signal)."""
    g = torch.Generator().manual_seed(seed)
    out = torch.zeros(batch, taxa, sites, dtype=torch.int64)
    for b in range(batch):
        seqs = [torch.randint(0, 4, (sites,), generator=g)]
        while len(seqs) < taxa:
            k = int(torch.randint(0, len(seqs), (1,), generator=g))
            parent = seqs.pop(k)
            for _ in range(2):
                child = parent.clone()
                mut = torch.rand(sites, generator=g) < rate * float(torch.rand(1, generator=g) + 0.25)
                child[mut] = torch.randint(0, 4, (int(mut.sum()),), generator=g)
                seqs.append(child)
        perm = torch.randperm(taxa, generator=g)
        m = torch.stack(seqs)[perm]
        m[torch.rand(taxa, sites, generator=g) < gap] = 4
        out[b] = m
    table = torch.tensor([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1], [1, 1, 1, 1]], dtype=torch.int8)
    return table[out]
