"""CPU: the host orchestration of `Argmax_inference(files_per_call=N)` - grouping alignments of equal shape into one rollout call,
per-file `.tre` output, file sharding over the ranks of a process group (gloo, world size 2).

The device model is replaced by a stand-in whose `rollout_fused` is the CPU oracle (test infrastructure; the product path has no
such fallback), so everything around the one device call - `load_pi_instance`, batch assembly, `reinforce_rollout`,
`PhyInferEnv.init_states / replay_merges`, Newick writing - is the shipped code and the trees can be compared with the Newick
strings of the executed reference (tests/golden)."""
import os
import shutil
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import GOLD

NAMES = ["t20x256_10", "t20x256_103", "t20x256_104", "t20x256_117", "t20x512_a"]      # four of one shape, one of another


class OracleAgent(torch.nn.Module):
    """Stand-in for PhyloATTN: same `rollout_fused` contract (merges int32 [B,R-1,2], selected log-probs [B,R-1], logits trace)."""

    def __init__(self):
        super().__init__()
        import nnj_oracle as O
        self.O, self.sd = O, O.init_state_dict(0)
        self.anchor = torch.nn.Parameter(torch.zeros(1))
        self.calls = []

    def rollout_fused(self, data, mask, gumbel=None, want_logits=False):
        B, R = data.shape[:2]
        self.calls.append((B, R, data.shape[2]))
        ref = self.O.rollout(self.sd, data, mask)
        total = sum(n * (n - 1) // 2 for n in range(R, 1, -1))
        trace = torch.zeros(B, total)
        flat = torch.cat(ref["logits"], dim=1)
        trace[:, :flat.shape[1]] = flat
        slp = torch.zeros(B, R - 1)
        slp[:, :R - 2] = ref["selected_log_ps"]
        return ref["merges"].to(torch.int32), slp, trace


def _phy_dir(tmp_path):
    d = tmp_path / "msas"
    d.mkdir()
    for n in NAMES:
        shutil.copyfile(os.path.join(GOLD, "msa", n + ".phy"), d / (n + ".phy"))
    return str(d)


def _read(path):
    with open(path) as f:
        return f.read().strip()


def test_batched_directory_mode_groups_by_shape_and_writes_reference_trees(tmp_path, golden):
    from neuralnj_b200 import Argmax_inference, inference_config
    src = _phy_dir(tmp_path)
    agent = OracleAgent()
    written = Argmax_inference(src, str(tmp_path / "out"), None, cfgs=inference_config(), device=torch.device("cpu"),
                               files_per_call=3, policy_network=agent)
    assert [os.path.basename(w) for w in written] == [n + ".tre" for n in sorted(NAMES)]
    # three 20 x 256 alignments in the first call, the fourth and the 20 x 512 one in calls of their own: nothing padded
    assert sorted(agent.calls) == [(1, 20, 256), (1, 20, 512), (3, 20, 256)]
    for n in NAMES:
        assert _read(tmp_path / "out" / (n + ".tre")) == golden(n).newick[0], n
    # the reference's per-file loop (files_per_call = 1) writes the same trees
    agent1 = OracleAgent()
    Argmax_inference(src, str(tmp_path / "out1"), None, cfgs=inference_config(), device=torch.device("cpu"), policy_network=agent1)
    assert len(agent1.calls) == len(NAMES) and all(c[0] == 1 for c in agent1.calls)
    for n in NAMES:
        assert _read(tmp_path / "out1" / (n + ".tre")) == _read(tmp_path / "out" / (n + ".tre"))


def _worker(rank, world, port, src, out_dir, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import conftest  # noqa: F401  (puts oracle/ on sys.path in the spawned process)
    from neuralnj_b200 import Argmax_inference, inference_config
    written = Argmax_inference(src, out_dir, None, cfgs=inference_config(), device=torch.device("cpu"), files_per_call=8,
                               policy_network=OracleAgent())
    q.put((rank, [os.path.basename(w) for w in written]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_share_the_file_list(tmp_path, golden):
    src = _phy_dir(tmp_path)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, src, str(tmp_path / "out"), q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    files = [n + ".tre" for n in sorted(NAMES)]
    assert got[0] == files[:3] and got[1] == files[3:]          # contiguous shares of the sorted list, sizes differ by at most one
    for n in NAMES:
        assert _read(tmp_path / "out" / (n + ".tre")) == golden(n).newick[0], n
