"""Table for DESIGN.md / profiles: multigrid-preconditioned vs diagonally preconditioned CG on the configs' inputs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import keff_cfd_b200 as K
from keff_cfd_b200 import api, structures as S
K.init(0)
G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
cases = []
img = K.structure_from_image(np.load(os.path.join(G, "img_00009.npy")))
cases.append(("00009.jpg 128^2", img[None]))
cases.append(("00009.jpg amp 2 (256^2)", K.structure_from_image(np.load(os.path.join(G, "img_00009.npy")), 2, 2)[None]))
for nm in ("schwarz163", "schwarz091"):
    cases.append((nm + " 128^3", np.unpackbits(np.load(os.path.join(G, nm + "_bits.npy"))).reshape(128, 128, 128)))
import torch
for n in (256, 512):
    cases.append((f"bernoulli {n}^3", api.bernoulli_volume_dev((n, n, n), 0.5, 1234)))
for name, m in cases:
    row = [name]
    for flags in (0, K.F_NO_MG):
        r = api.solve3d(m, api.default_params(flags=flags), want_T=False)
        r = api.solve3d(m, api.default_params(flags=flags), want_T=False)
        row.append(f"{r['iters']} it / {r['gpu_ms']:.2f} ms / keff {r['keff']:.10f}")
    print(" | ".join(row), flush=True)
