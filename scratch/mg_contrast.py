"""Robustness of the multigrid preconditioner against the conductivity ratio."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import keff_cfd_b200 as K
from keff_cfd_b200 import api, structures as S
K.init(0)
G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
vol = api.bernoulli_volume_dev((128, 128, 128), 0.5, 1234)
sch = np.unpackbits(np.load(os.path.join(G, "schwarz163_bits.npy"))).reshape(128, 128, 128)
img = K.structure_from_image(np.load(os.path.join(G, "img_00009.npy")))[None]
for name, m in (("bernoulli128", vol), ("schwarz163", sch), ("img00009", img)):
    for ks in (10.0, 100.0, 1e3, 1e4, 1e6):
        out = []
        for flags in (0, K.F_NO_MG):
            r = api.solve3d(m, api.default_params(ks=ks, flags=flags, max_iter=20000), want_T=False)
            out.append(f"{r['iters']} it {r['gpu_ms']:.1f} ms st {r['status']} keff {r['keff']:.8g}")
        print(name, "ks", ks, "| mg:", out[0], "| jacobi:", out[1], flush=True)
