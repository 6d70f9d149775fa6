#!/bin/bash
echo "== default"; timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "onchip_mixed or contrast" 2>&1 | tail -1
echo "== delta 1e-4"; KEFFB200_BATCH_DELTA=1e-4 timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "onchip_mixed or contrast" 2>&1 | tail -1
echo "== omega .85"; KEFFB200_BATCH_OMEGA=0.85 timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "onchip_mixed or contrast" 2>&1 | tail -1
cat > /tmp/t.py <<'PY'
import sys; sys.path.insert(0, '.')
import numpy as np, keff_cfd_b200 as K
from keff_cfd_b200 import structures as S
imgs = S.disc_images(8)
p = dict(ks=400.0, kf=0.026, tl=300.0, tr=290.0)
a = K.solve2d_batch(imgs, K.default_params(**p))
b = K.solve2d_batch(imgs, K.default_params(flags=K.F_BATCH_GLOBAL, **p))
print("mg  ", a["status"], a["iters"], a["rel_residual"])
print("fp64", b["status"], b["iters"])
print("gap ", np.abs(a["keff"] - b["keff"]) / b["keff"])
PY
python /tmp/t.py
KEFFB200_BATCH_DELTA=1e-4 python /tmp/t.py
