import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import keff_cfd_b200 as K
from keff_cfd_b200 import api, structures as S
n = int(sys.argv[1]) if len(sys.argv) > 1 else 296
flags = int(sys.argv[2]) if len(sys.argv) > 2 else 0
K.init(0)
imgs = torch.from_numpy(S.disc_images(n)).cuda()
p = api.default_params(flags=flags)
for _ in range(2):
    r = api.solve2d_batch(imgs, p)
its = np.array([x["iters"] for x in r])
print(f"{n} images flags {flags}: {r[0]['gpu_ms']:.2f} ms -> {n / r[0]['gpu_ms'] * 1e3:.0f} solves/s; mean iters {its.mean():.0f}; "
      f"cycles/iter/SM ~ {r[0]['gpu_ms'] * 1e-3 * 1.965e9 * min(n,148) / its.sum():.0f}; all ok {all(x['status'] == 0 for x in r)}")
