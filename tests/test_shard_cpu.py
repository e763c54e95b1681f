"""CPU, world_size 2 (gloo): the N>1 plumbing of bench.py / shard.py — alignment sharding, max-over-ranks timing, gather."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from neuralnj_b200.shard import gather_merges, shard_bounds, sharded_search


def _worker(rank, world, port, total, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_bounds(total, rank, world)
    # stand-in for the device rollout: every alignment's "merge list" encodes its global index
    local = torch.arange(lo, hi, dtype=torch.int32).view(-1, 1, 1).expand(-1, 3, 2).contiguous()
    full = gather_merges(local, total, rank, world)
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    # Search mode: rank r "finds" a tree with score -10 + r from its own seed; every rank must end up with rank 1's tree
    res = sharded_search(lambda seed: dict(the_best_score=-10.0 + (seed - 40), the_best_tree=f"(tree_of_seed_{seed});", distinct_topologies=seed), seed=40)
    assert res["the_best_tree"] == "(tree_of_seed_41);" and res["best_rank"] == 1 and res["distinct_topologies_per_rank"] == [40, 41]
    if rank == 0:
        out.put((full[:, 0, 0].tolist(), float(t)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_gather():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    total = 7
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    got, tmax = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got == list(range(total)) and tmax == 2.0


def test_shard_bounds_cover_everything():
    for total in (1, 7, 512, 513):
        for world in (1, 2, 4, 8):
            spans = [shard_bounds(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_sharded_search_single_process():
    res = sharded_search(lambda seed: dict(the_best_score=-3.5, the_best_tree="(a,b,c);", distinct_topologies=2), seed=9)
    assert res["the_best_tree"] == "(a,b,c);" and res["best_rank"] == 0
