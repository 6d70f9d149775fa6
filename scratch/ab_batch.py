"""A/B of the 2D batch kernels: on-chip multigrid (default) against the two-level additive kernel (flag 16).
usage: ab_batch.py [n_images] ; env KEFFB200_BATCH_OMEGA / KEFFB200_BATCH_DELTA tune the multigrid kernel."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import keff_cfd_b200 as K
from keff_cfd_b200 import api, structures as S
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1480
K.init(0)
worst = 0.0
for shape in [(16, 16), (32, 16), (48, 80), (64, 64), (128, 64), (96, 128), (128, 128)]:
    ny, nx = shape
    rng = np.random.default_rng(ny * 1000 + nx)
    m = (rng.random((5, ny, nx)) < 0.45).astype(np.uint8)
    m[4] = S.disc_images(1, 128)[0][:ny, :nx]
    a = api.solve2d_batch(torch.from_numpy(m).cuda(), api.default_params())
    b = api.solve2d_batch(torch.from_numpy(m).cuda(), api.default_params(flags=16))
    e = np.abs(a["keff"] - b["keff"]) / b["keff"]
    worst = max(worst, e.max())
    print(shape, "iters mg", a["iters"], "add", b["iters"], "status", a["status"], "relres", a["rel_residual"].max(), "keff gap", e.max(), flush=True)
imgs = torch.from_numpy(S.disc_images(n)).cuda()
for name, flags in (("mg", 0), ("additive", 16)):
    p = api.default_params(flags=flags)
    for _ in range(3):
        r = api.solve2d_batch(imgs, p)
    print(f"{name}: {n} images {r['gpu_ms'][0]:.2f} ms -> {n / r['gpu_ms'][0] * 1e3:.0f} solves/s; mean iters {r['iters'].mean():.1f} max {r['iters'].max()}; "
          f"all ok {bool((r['status'] == 0).all())}; max relres {r['rel_residual'].max():.2e}; mean keff {r['keff'].mean():.10f}", flush=True)
    if flags == 0: k0 = r["keff"].copy()
    else: print("max keff gap mg vs additive", (np.abs(k0 - r["keff"]) / r["keff"]).max())
print("worst small-shape gap", worst)
