"""GPU: encoder parity vs the CPU oracle for every NNJ_ENC_TC stage mask (one subprocess per mask; the mask is read once per process)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, torch
sys.path.insert(0, %r); sys.path.insert(0, %r + "/oracle")
import nnj_oracle as O
from neuralnj_b200 import PhyloATTN, inference_config
for nl in (1, 6):
    cfg = inference_config(); cfg.model.num_enc_layers = nl
    torch.manual_seed(0); m = PhyloATTN(cfg, precision="bf16x3").cuda().eval()
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    for (B, R, L, pad) in ((2, 20, 256, 0), (1, 50, 128, 0), (2, 7, 96, 24), (1, 100, 64, 0), (1, 130, 64, 0)):
        data = O.evolved_msa(B, R, L, seed=5)
        mask = torch.zeros(B, L, dtype=torch.bool)
        if pad:
            mask[:, L - pad:] = True; data[:, :, L - pad:, :] = 0
        want = O.encode(sd, data, mask)
        got = m.encode_zxr(data.cuda(), mask.cuda()).cpu()
        err = float((got - want).abs().max())
        print(f"  layers={nl} B={B} R={R} L={L} pad={pad}: max|err|={err:.3e} {'OK' if err < 2e-4 else 'FAIL'}", flush=True)
''' % (ROOT, ROOT)
for mk in [int(v) for v in (sys.argv[1:] or ["0", "1", "4", "2", "7"])]:
    print(f"== NNJ_ENC_TC={mk}", flush=True)
    env = dict(os.environ, NNJ_ENC_TC=str(mk))
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=600)
    print(r.stdout, end="")
    if r.returncode != 0:
        print("  child failed:", r.stderr[-1500:])
