"""B200: `Argmax_inference(files_per_call=N)` - alignments of equal shape stacked into one device rollout - writes the trees of the
executed reference (the host orchestration is covered on the CPU by tests/test_entry_batched_cpu.py; this is the same call on the
CUDA path in the default precision).  Sorted last on purpose: the mode was added after the round's last GPU session."""
import os
import shutil

import pytest
import torch

from conftest import GOLD

pytestmark = pytest.mark.gpu


def test_batched_argmax_inference_writes_reference_trees(tmp_path, golden):
    from neuralnj_b200 import Argmax_inference, inference_config
    names = ["t20x256_10", "t20x256_120", "t20x256_103", "ex50x1024_73"]
    src = tmp_path / "msas"
    src.mkdir()
    for n in names:
        shutil.copyfile(os.path.join(GOLD, "msa", n + ".phy"), src / (n + ".phy"))
    torch.manual_seed(0)
    written = Argmax_inference(str(src), str(tmp_path / "out"), None, cfgs=inference_config(), files_per_call=4)
    assert [os.path.basename(w) for w in written] == [n + ".tre" for n in sorted(names)]
    for n in names:
        with open(tmp_path / "out" / (n + ".tre")) as f:
            assert f.read().strip() == golden(n).newick[0], n
