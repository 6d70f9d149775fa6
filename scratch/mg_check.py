"""First-light / timing driver of the multigrid path: iteration counts and device time, with and without MG."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import keff_cfd_b200 as K
from keff_cfd_b200 import api, structures as S
K.init(0)
sizes = [int(a) for a in sys.argv[1:]] or [32, 64, 128, 256]
for n in sizes:
    vol = api.bernoulli_volume_dev((n, n, n), 0.5, 1234)
    for flags, name in ((0, "mg"), (K.F_NO_MG, "jacobi")):
        if name == "jacobi" and n > 256:
            continue
        r = api.solve3d(vol, api.default_params(flags=flags), want_T=False)
        r2 = api.solve3d(vol, api.default_params(flags=flags), want_T=False)
        print(f"{n}^3 {name}: iters {r2['iters']} keff {r2['keff']:.10f} relres {r2['rel_residual']:.2e} status {r2['status']} "
              f"gpu_ms {r2['gpu_ms']:.2f} (first call {r['gpu_ms']:.2f}) ms/iter {r2['gpu_ms']/max(r2['iters'],1):.3f}", flush=True)
    del vol
img = S.disc_images(1, 128)[0]
for flags, name in ((0, "mg"), (K.F_NO_MG, "jacobi")):
    r = K.solve2d(img, K.default_params(flags=flags), want_T=False)
    print("disc128", name, r["iters"], r["keff"], r["gpu_ms"])
