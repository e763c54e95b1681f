import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import llh_oracle as O
from neuralnj_b200 import likelihood as LH
R, L, B = 50, 1024, 8
rng = np.random.default_rng(9)
gen = O.Model(rates6=(1.2, 3.1, 0.8, 1.1, 4.2, 1.0), freqs=(0.3, 0.2, 0.2, 0.3), alpha=0.7, pinv=0.15)
ch0, bl = O.random_tree(R, rng)
tips = O.simulate(ch0, bl, R, L, gen, rng, gap_frac=0.05)
ch = np.stack([O.random_tree(R, rng)[0] for _ in range(B)])
pats, w = LH.compress_patterns(tips)
eng = LH.TreeLikelihood(pats, w)
sm = LH.SubstModel("GTR+I+G", LH.empirical_freqs(tips), B)
brl = np.full((B, 2 * R - 2), 0.1)
eng.loglik(ch, brl, sm)                                   # launch 0: eval
eng.optimize_branches(ch, brl, sm, max_passes=1)          # launch 1: one sweep
eng.optimize_all(ch, brl, sm, max_rounds=1)               # launch 2: branch opt + one round of model search
torch.cuda.synchronize()
