"""A/B of the stencil kernels: correctness of the current build's TMA path against the plain global-load kernel
(F_NO_TMA) on assorted shapes, then the 512^3 sweep time.  KEFFB200_STENCIL_WS=1 selects the warp-specialised ring."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import keff_cfd_b200 as K
from keff_cfd_b200 import api, structures as S
K.init(0)
rng = np.random.default_rng(5)
worst = 0.0
for shape in [(1, 2, 16), (1, 16, 16), (1, 100, 128), (3, 5, 32), (7, 20, 48), (33, 17, 80), (64, 64, 64), (130, 70, 144)]:
    m = S.bernoulli_volume(shape, 0.5, 7)
    x = rng.standard_normal(shape)
    y1, b1 = api.apply_operator(m, x, api.default_params(), want_b=True)
    y0, b0 = api.apply_operator(m, x, api.default_params(flags=1), want_b=True)
    e = max(np.abs(y1 - y0).max() / np.abs(y0).max(), np.abs(b1 - b0).max())
    worst = max(worst, e)
    print(shape, "rel err vs plain kernel", e)
assert worst < 1e-13, worst
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
vol = api.bernoulli_volume_dev((n, n, n), 0.5, 1234)
for rep in range(3):
    ms = api.bench_apply(vol, api.default_params(), reps=20)
    print(f"apply {n}^3 [{'ws' if os.environ.get('KEFFB200_STENCIL_WS') else 'sync'}]: {ms:.4f} ms, {n**3/ms/1e6:.1f} Gcell/s, {16*n**3/ms/1e6:.0f} GB/s = {16*n**3/ms/1e6/6538.9:.3f} of measured peak")
r = api.solve3d(vol, api.default_params(), want_T=False)
r = api.solve3d(vol, api.default_params(), want_T=False)
print("solve", r["iters"], r["gpu_ms"], "ms keff", r["keff"], "relres", r["rel_residual"], "status", r["status"])
