import sys, time, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import llh_oracle as O
from neuralnj_b200 import likelihood as LH
R, L = 50, 1024
rng = np.random.default_rng(9)
gen = O.Model(rates6=(1.2, 3.1, 0.8, 1.1, 4.2, 1.0), freqs=(0.3, 0.2, 0.2, 0.3), alpha=0.7, pinv=0.15)
ch0, bl = O.random_tree(R, rng)
tips = O.simulate(ch0, bl, R, L, gen, rng, gap_frac=0.05)
for B in (1, 16, 64, 128):
    ch = np.stack([O.random_tree(R, rng)[0] for _ in range(B)])
    pats, w = LH.compress_patterns(tips)
    eng = LH.TreeLikelihood(pats, w)
    sm = LH.SubstModel("GTR+I+G", LH.empirical_freqs(tips), B)
    brl = np.full((B, 2 * R - 2), 0.1)
    eng.loglik(ch, brl, sm)
    def tm(f, n=3):
        torch.cuda.synchronize(); t0 = time.time()
        for _ in range(n): r = f()
        torch.cuda.synchronize(); return (time.time() - t0) / n, r
    t_eval, _ = tm(lambda: eng.loglik(ch, brl, sm), 5)
    t_br, (t1, _, ll1) = tm(lambda: eng.optimize_branches(ch, brl, sm), 2)
    t_br1, _ = tm(lambda: eng.optimize_branches(ch, brl, sm, max_passes=1), 2)
    out = {}
    for rounds in (1, 10):
        sm2 = LH.SubstModel("GTR+I+G", LH.empirical_freqs(tips), B)
        t_all, (_, _, ll2) = tm(lambda: eng.optimize_all(ch, brl, sm2, max_rounds=rounds), 1)
        out[rounds] = (round(t_all, 3), float(ll2.mean()))
    print(f"B={B}: eval {t_eval*1e3:.2f} ms, 1 sweep {t_br1*1e3:.1f} ms, branch opt {t_br*1e3:.1f} ms, optimise_all {out}", flush=True)
