"""Small driver for ncu: the 512^3 stencil kernel (and optionally a few CG iterations)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import keff_cfd_b200 as K
from keff_cfd_b200 import api
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 0
K.init(0)
vol = api.bernoulli_volume_dev((n, n, n), 0.5, 1234)
ms = api.bench_apply(vol, api.default_params(), reps=5)
print(f"apply {n}^3: {ms:.4f} ms, {n**3/ms/1e6:.1f} Gcell/s, {16*n**3/ms/1e6:.0f} GB/s")
if iters:
    r = api.solve3d(vol, api.default_params(max_iter=iters, check_every=iters), want_T=False)
    print("solve", r["iters"], r["gpu_ms"], "ms ->", r["gpu_ms"] / max(r["iters"], 1), "ms/iter")
