"""CPU: host-side mirror of the reference interface (no GPU compute) and the C-ABI surface."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import GOLD, ROOT

import neuralnj_b200 as nnj
from neuralnj_b200 import _lib
from neuralnj_b200.environment import PhyInferEnv, format_rtree, format_rtree_topology
from neuralnj_b200.treeutil import bipartitions, normalized_rf, rf_distance, treestr_to_tuples


def test_library_builds_loads_and_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    L = _lib.lib()
    hdr = open(os.path.join(ROOT, "include", "nnj.h")).read()
    declared = set(re.findall(r"\b(nnj_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for sym in declared:
        assert hasattr(raw, sym), sym
    assert L.nnj_abi_version() == 1
    assert L.nnj_profile_classes() >= 16 and L.nnj_profile_name(11) == b"pair_score"
    # argument validation happens before any CUDA call
    assert L.nnj_encode(None, None, None, 1, 2, 8, None, None, 0, None) == -1
    assert b"bad arguments" in L.nnj_last_error()


def test_build_is_keyed_by_source_contents_not_file_times():
    """A snapshot copy (the GPU box) does not keep file times: build() must recognise an up-to-date libnnj.so by the hash of
    flags + sources + headers recorded beside it, and concurrent callers (the ranks of one torchrun job) must not rebuild it."""
    import subprocess
    import sys
    _lib.build()
    flags = " ".join(_lib.NVCC_FLAGS + os.environ.get("NNJ_EXTRA_NVCC_FLAGS", "").split())
    assert open(_lib.HASH_TAG).read() == _lib._source_hash(flags)
    before = os.stat(_lib.LIB_PATH).st_mtime_ns
    src = os.path.join(os.path.dirname(_lib.LIB_PATH), "csrc", "nnj_api.cu")
    st = os.stat(src)
    try:
        os.utime(src)                                    # newer than the library, same bytes
        code = "from neuralnj_b200 import _lib; _lib.build(); print(_lib.lib().nnj_abi_version())"
        procs = [subprocess.Popen([sys.executable, "-c", code], cwd=ROOT, stdout=subprocess.PIPE, text=True) for _ in range(2)]
        assert [p.communicate(timeout=600)[0].strip() for p in procs] == ["1", "1"]
    finally:
        os.utime(src, ns=(st.st_atime_ns, st.st_mtime_ns))
    assert os.stat(_lib.LIB_PATH).st_mtime_ns == before


def test_sm100a_cubin_is_embedded():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_model_mirror_has_reference_state_dict_layout(sd0):
    torch.manual_seed(0)
    m = nnj.PhyloATTN(nnj.inference_config())
    sd = m.state_dict()
    assert list(sd.keys()) == list(sd0.keys())
    assert all(torch.equal(sd[k], sd0[k]) for k in sd)          # same RNG consumption order as the reference ctor
    assert sum(p.numel() for p in m.parameters()) == 425857
    m.load_state_dict({k: v.clone() for k, v in sd0.items()})    # a reference checkpoint loads unchanged
    with pytest.raises(nnj.NnjError):
        m.encode_zxr(torch.zeros(1, 4, 8, 4, dtype=torch.int8), torch.zeros(1, 8, dtype=torch.bool))   # CPU: no fallback
    with pytest.raises(NotImplementedError):
        m(torch.zeros(1))


def test_config_defaults_and_yaml_merge(tmp_path):
    c = nnj.empty_config()
    assert (c.model.patch_size, c.model.embed_dim, c.env.batch_size) == (4, 32, 8)      # utils.py:44-51
    p = tmp_path / "c.yaml"
    p.write_text("env:\n  batch_size: 1\nmodel:\n  patch_size: 1\n  embed_dim: 64\n  num_enc_heads: 8\n  num_enc_layers: 6\n")
    c.merge_from_file(str(p))
    i = nnj.inference_config()
    assert dict(c.model) == dict(i.model) and c.env.batch_size == 1 and c.model.vocab_size == 4


def test_load_pi_instance_matches_reference_loader(golden):
    for case in ("t20x256_103", "t100x256_a", "ex50x1024_71"):
        g = golden(case)
        b = nnj.load_pi_instance(os.path.join(GOLD, "msa", case + ".phy"))
        assert torch.equal(b["data"], g.data) and b["data"].dtype == torch.int8
        assert b["seq_keys"] == g.seq_keys and bool((b["seq_weights"] == 1).all())
        assert b["taxa_nums"] == [g.data.shape[1]] and b["seq_lens"] == [g.data.shape[2]]


def test_loaders_edge_cases(tmp_path):
    p = tmp_path / "x.phy"          # interleaved, lower case, unknown symbols, shuffled taxa
    p.write_text("3 8\nt2 acgt\nt3 ----\nt1 NNKK\n\nTTTT\nAC?T\nGGGG\n")
    b = nnj.load_pi_instance(str(p))
    assert b["seq_keys"] == [["t1", "t2", "t3"]]
    assert b["seqs"][0] == ["NN--GGGG", "ACGTTTTT", "----AC-T"]
    d = b["data"][0]
    assert d[0, 0].tolist() == [1, 1, 1, 1] and d[1, 1].tolist() == [0, 1, 0, 0] and d[2, 6].tolist() == [1, 1, 1, 1]
    f = tmp_path / "y.fasta"         # ragged rows: columns past the shortest row become '*' padding with weight 0
    f.write_text(">a\nACGTAC\n>b\nACG\n")
    b = nnj.load_pi_instance(str(f))
    assert b["seq_weights"][0].tolist() == [1, 1, 1, 0, 0, 0] and int(b["data"][0, :, 3:].abs().sum()) == 0
    bad = tmp_path / "z.phy"
    bad.write_text("2 4\nhomo_sapiens ACGT\nmus ACGT\n")
    with pytest.raises(ValueError):
        nnj.load_pi_instance(str(bad))
    with pytest.raises(ValueError):
        nnj.load_pi_instance(str(tmp_path / "x.txt"))


def test_env_replay_reproduces_reference_newick(golden):
    cfg = nnj.inference_config()
    for case in ("t20x256_10", "batch2_20x256", "tiny_3x64", "t100x256_a"):
        g = golden(case)
        B, R = g.data.shape[:2]
        env = PhyInferEnv(cfg, torch.device("cpu"))
        env.init_states([["A"] * R] * B, g.seq_keys, g.data)
        assert env.tree_pairs_dict[R][:2] == [(0, 1), (0, 2)] and len(env.tree_pairs_dict[R]) == R * (R - 1) // 2
        env.replay_merges(g.merges)
        scores, rt, ut, best = env.evaluate_loglikelihood()
        assert [s.subtrees[0].utree_op_str for s in env.states] == g.newick
        assert best == g.newick[0] and float(scores[0]) == -111111
        trees, sc = env.dump_end_trees()
        assert sc == [-111111] * B and trees[0].topo_repr.count(",") == R - 1
        assert rt == treestr_to_tuples(g.newick[0])
        all_scores, rts, uts, strs = env.evaluate_loglikelihood(get_all_tree=True)
        assert strs == g.newick


def test_env_step_host_half_with_mean_agent():
    """env.step(agent=None) averages the two embeddings (environment.py:826) and re-indexes slot i <- new, j removed."""
    cfg = nnj.inference_config()
    env = PhyInferEnv(cfg, torch.device("cpu"))
    keys = [[f"t{i + 1}" for i in range(4)]]
    env.init_states([["A"] * 4], keys, None)
    env.state_tensor = torch.arange(4.0).view(1, 4, 1, 1).expand(1, 4, 2, 3).clone()
    a = env.action_indices_dict[4][(1, 3)]
    assert env.step([a], [(None, None)], branch_optimize=False, agent=None) is False
    assert env.state_tensor[0, :, 0, 0].tolist() == [0.0, 2.0, 2.0]
    assert env.get_current_trees()[0] == ["0;", "(1, 3);", "2;"] or env.get_current_trees()[0][1].startswith("(1, 3)")
    assert env.step([env.action_indices_dict[3][(0, 2)]], [(None, None)], branch_optimize=False, agent=None) is False
    with pytest.raises(RuntimeError):
        env.step([0], [(None, None)], branch_optimize=True, agent=None)      # RAxML-NG scoring is out of scope


def test_newick_helpers():
    t = treestr_to_tuples("((a:0.1, b:0.2):0.3, c:0.4);")
    assert t == (("a", 0.1, "b", 0.2), 0.3, "c", 0.4)
    x = "((t1:1, t2:1):1, (t3:1, t4:1):1, t5:1);"
    y = "((t1:1, t3:1):1, (t2:1, t4:1):1, t5:1);"
    assert rf_distance(x, x) == 0 and rf_distance(x, y) == 4 and normalized_rf(x, y) == 1.0
    assert len(bipartitions(x)) == 2


def test_rollout_driver_refuses_training_mode():
    with pytest.raises(NotImplementedError):
        nnj.reinforce_rollout({}, None, None, None, eval=False)


def test_library_path_override_fails_loudly(tmp_path):
    """NNJ_LIB_PATH points the binding at another build of libnnj (A/B runs); a missing file raises, nothing falls back."""
    import subprocess
    import sys
    code = ("import os, sys; sys.path.insert(0, %r)\n"
            "from neuralnj_b200 import _lib\n"
            "try:\n    _lib.lib()\nexcept _lib.NnjError as e:\n    print('NnjError'); sys.exit(0)\nsys.exit(1)\n") % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, NNJ_LIB_PATH=str(tmp_path / "absent.so"))
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "NnjError" in r.stdout, r.stderr
