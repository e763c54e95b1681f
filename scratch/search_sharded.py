"""Config 3 over several GPUs (BASELINE: "Search-mode inference, 50 taxa x 1024 sites, batched on 8 x B200"): every rank encodes the same
alignment once, samples its own rollouts from a rank-specific seed, scores the distinct topologies with the GPU likelihood and the best tree over
all ranks wins (neuralnj_b200.shard.sharded_search; only (score, Newick) pairs cross the ranks).
   torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scratch/search_sharded.py [episodes] [batch]"""
import json, os, sys, time
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neuralnj_b200 import PhyInferEnv, PhyloATTN, RL_Search, inference_config
from neuralnj_b200.shard import sharded_search

episodes = int(sys.argv[1]) if len(sys.argv) > 1 else 2
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 50
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
cfgs = inference_config()
cfgs.env.batch_size = batch
cfgs.num_episodes = episodes
torch.manual_seed(0)
model = PhyloATTN(cfgs, precision="bf16x3").to(dev).eval()
env = PhyInferEnv(cfgs, dev)
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "msa", "ex50x1024_73.phy")
def one(seed):
    return RL_Search(cfgs, path, model, env, stop_step=episodes, generator=torch.Generator(device=dev).manual_seed(seed))
one(1000 + rank)          # warm-up (library load, workspaces)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.time()
res = sharded_search(one, seed=7)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
dt = time.time() - t0
if rank == 0:
    print(json.dumps({"what": "Search mode, 50 x 1024 (examples/len1024taxa50 ..._73.phy), sampled rollouts + GPU likelihood scoring (GTR+I+G, model optimised), sharded by trajectories",
                      "n_gpus": world, "rollouts_per_rank": episodes * batch, "rollouts_total": episodes * batch * world, "seconds": round(dt, 2),
                      "rollouts_per_s": round(episodes * batch * world / dt, 1), "best_llh": res["the_best_score"], "best_rank": res["best_rank"],
                      "distinct_topologies_per_rank": res.get("distinct_topologies_per_rank", [res.get("distinct_topologies")])}))
if world > 1:
    dist.destroy_process_group()
