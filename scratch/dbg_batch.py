import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import keff_cfd_b200 as K
from keff_cfd_b200 import api, structures as S
start, count = int(sys.argv[1]), int(sys.argv[2])
dev = int(sys.argv[3]) if len(sys.argv) > 3 else 0
torch.cuda.set_device(dev)
K.init(dev)
imgs = torch.from_numpy(S.disc_images(count, 128, start=start)).cuda()
r = api.solve2d_batch(imgs, api.default_params())
print("ok", start, count, "dev", dev, "iters", r["iters"].mean(), r["iters"].max(), "status", np.unique(r["status"]), "relres", r["rel_residual"].max(), flush=True)
