"""Multi-GPU plumbing.  Alignments are independent: a batch is cut into contiguous
per-rank shards, every rank runs the whole hot path on its shard, and only the tiny merge lists are gathered.
No collective sits on the data path; `torch.distributed` (NCCL on GPUs, gloo in the CPU tests) is used for the
barrier, the max-over-ranks timing and this gather."""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_bounds(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [lo, hi) of `total` alignments for `rank` (sizes differ by at most one)."""
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_merges(local: torch.Tensor, total: int, rank: int, world: int) -> torch.Tensor:
    """All ranks contribute their [n_local, R-1, 2] int32 merge lists; every rank gets [total, R-1, 2] in input order."""
    if world == 1:
        return local
    sizes = [shard_bounds(total, r, world) for r in range(world)]
    pad = max(h - l for l, h in sizes)
    buf = torch.zeros((pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    buf[: local.shape[0]] = local
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)
    return torch.cat([p[: h - l] for p, (l, h) in zip(parts, sizes)], dim=0)


def sharded_rollout(model, data: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """Argmax rollouts of a whole batch, sharded by alignment over the ranks of the default process group."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    lo, hi = shard_bounds(data.shape[0], rank, world)
    merges, _, _ = model.rollout_fused(data[lo:hi], mask[lo:hi])
    return gather_merges(merges, data.shape[0], rank, world)


def sharded_search(search_fn, seed: int = 0) -> dict:
    """Search mode (NeuralNJ-MC) over the ranks of the default process group - BASELINE config 3, "batched on 8 x B200".
    Every rank runs `search_fn(rank_seed)`: its own sampled rollouts of the SAME alignment (shared encoder pass, on-device
    Gumbel-max, likelihood scoring on its GPU; e.g. `lambda s: RL_Search(cfgs, path, model, env, generator=make_gen(s))`) with a
    rank-specific seed, so the ranks explore different trajectories.  Only (score, Newick string) pairs are exchanged; the
    best-scoring tree over all ranks is returned on every rank (ties go to the lowest rank: deterministic)."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    res = dict(search_fn(seed + rank))
    if world == 1:
        res["best_rank"] = 0
        return res
    mine = (float(res["the_best_score"]), res["the_best_tree"], int(res.get("distinct_topologies", 0)))
    every = [None] * world
    dist.all_gather_object(every, mine)
    best = max(range(world), key=lambda r: (every[r][0], -r))
    res.update(the_best_score=every[best][0], the_best_tree=every[best][1], best_rank=best,
               distinct_topologies_per_rank=[e[2] for e in every])
    return res
