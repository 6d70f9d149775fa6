"""One process, every visible GPU (keffb200_init_multi) against one GPU: image batch and z-slabbed volume."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import keff_cfd_b200 as K
from keff_cfd_b200 import _lib, api, structures as S

imgs = S.disc_images(600, n=64)
vol = S.bernoulli_volume((96, 64, 64), 0.5, 4321)
K.init(0)
b1 = api.solve2d_batch(imgs)
v1 = api.solve3d(vol, want_q=True)
K.shutdown()
n = _lib.init_multi(0)
assert n >= 2, n
_lib.lib().keffb200_set_slab_min_cells(0)           # slab even this small volume
l0 = _lib.lib().keffb200_launch_count()
bn = api.solve2d_batch(imgs)
vn = api.solve3d(vol, want_q=True)
assert _lib.lib().keffb200_launch_count() > l0
assert (bn["status"] == 0).all() and np.array_equal(bn["iters"], b1["iters"])
assert np.abs(bn["keff"] - b1["keff"]).max() == 0.0, "a batch image must not depend on the GPU it ran on"
print("batch over", n, "GPUs: identical to one GPU;", "volume:", vn["keff"], v1["keff"], vn["iters"], v1["iters"])
assert vn["status"] == 0 and abs(vn["keff"] - v1["keff"]) / v1["keff"] < 1e-7
assert np.abs(vn["T"] - v1["T"]).max() < 1e-6
assert np.abs(vn["qL"] - v1["qL"]).max() < 1e-6 and np.abs(vn["qR"] - v1["qR"]).max() < 1e-6
# warm start through the slab path
vw = api.solve3d(vol, T_init=v1["T"])
assert vw["status"] == 0 and vw["iters"] <= 1 and abs(vw["keff"] - v1["keff"]) / v1["keff"] < 1e-7
K.shutdown()
print("multi ok")
