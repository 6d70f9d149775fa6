"""Stress of k_batch2d_mg against the all-FP64 kernel: every side combination in {16..128 step 16}^2, several
conductivity ratios, random and disc structures.  Prints the worst disagreement and any non-converged record."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import keff_cfd_b200 as K
from keff_cfd_b200 import structures as S
K.init(0)
rng = np.random.default_rng(3)
worst, bad, n_solved, max_it = 0.0, [], 0, 0
discs = S.disc_images(8, 128)
for ny in range(16, 129, 16):
    for nx in range(16, 129, 16):
        imgs = np.concatenate([(rng.random((3, ny, nx)) < p).astype(np.uint8) for p in (0.2, 0.5, 0.8)] + [discs[:3, :ny, :nx]])
        for ks, kf in ((10.0, 1.0), (1.0, 25.0), (300.0, 1.0), (1.0, 1.0)):
            p = dict(ks=ks, kf=kf, tl=1.0, tr=3.0)
            a = K.solve2d_batch(imgs, K.default_params(**p))
            b = K.solve2d_batch(imgs, K.default_params(flags=K.F_BATCH_GLOBAL, **p))
            n_solved += len(imgs)
            max_it = max(max_it, int(a["iters"].max()))
            gap = np.abs(a["keff"] - b["keff"]) / np.abs(b["keff"])
            worst = max(worst, float(gap.max()))
            if (a["status"] != 0).any() or (a["rel_residual"] > 1e-8).any() or gap.max() > 1e-5:
                bad.append((ny, nx, ks, kf, a["status"].tolist(), float(gap.max())))
print("solved", n_solved, "worst keff gap vs fp64 kernel", worst, "max iters", max_it, "bad", bad[:10], len(bad))
