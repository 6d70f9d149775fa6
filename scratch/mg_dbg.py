import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import keff_cfd_b200 as K
from keff_cfd_b200 import structures as S
K.init(0)
shape = (1, 128, 128)
m = S.bernoulli_volume(shape, 0.5, 11)
a = np.random.default_rng(0).standard_normal(shape)
z = K.precondition(m, a)
print("ok", float((z * a).sum()))
