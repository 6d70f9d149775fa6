"""Small driver for ncu: a few multigrid-preconditioned CG iterations at n^3."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import keff_cfd_b200 as K
from keff_cfd_b200 import api
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
K.init(0)
vol = api.bernoulli_volume_dev((n, n, n), 0.5, 1234)
if os.environ.get("PROF_WARM", "1") == "1" and iters > 100:
    api.solve3d(vol, api.default_params(max_iter=iters), want_T=False)   # first call allocates the workspace
r = api.solve3d(vol, api.default_params(max_iter=iters), want_T=False)
print("solve", r["iters"], r["gpu_ms"], "ms ->", r["gpu_ms"] / max(r["iters"], 1), "ms/iter", "keff", repr(r["keff"]), "relres", r["rel_residual"], "status", r["status"])
