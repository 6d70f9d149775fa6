"""Iteration counts of candidate preconditioners for the on-chip 2D batch kernel (numpy, CPU only)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from proto_mg_general import Level, fine_level, restrict, prolong, dense_inverse, pcg
from keff_cfd_b200 import structures as S

def build(solid, jump=1, half=False):
    L0 = fine_level(solid)
    levels, flags = [L0], []
    while np.prod(levels[-1].shape) > 64:
        L = levels[-1]; fl = None
        for _ in range(jump):
            if np.prod(L.shape) <= 64: break
            L, f = L.coarsen(); fl = f if fl is None else fl
        if half:
            h = lambda a: a.astype(np.float16).astype(float)
            L = Level(h(L.cx), h(L.cy), h(L.cz), h(L.wl), h(L.wr), L.shape)
        levels.append(L); flags.append((fl, jump))
    return levels, flags, dense_inverse(levels[-1])

def R(r, fj, Lf, Lc):
    f, j = fj
    for _ in range(j):
        if r.shape == Lc.shape: break
        r = restrict(r, f)
    return r
def P(e, fj, shape):
    f, j = fj
    shapes = [shape]
    s = shape
    for _ in range(j):
        s = tuple((n + 1) // 2 if ff else n for n, ff in zip(s, f)); shapes.append(s)
        if s == e.shape: break
    for s in reversed(shapes[:-1]):
        e = prolong(e, f, s)
    return e

def vcycle(levels, flags, l, b, w, dense):
    L = levels[l]
    if l == len(levels) - 1:
        return (dense @ b.ravel()).reshape(b.shape)
    x = np.zeros_like(b)
    for wk in w:
        x = x + wk * (b - L.apply(x)) / L.d
    e = vcycle(levels, flags, l + 1, R(b - L.apply(x), flags[l], L, levels[l + 1]), w, dense)
    x = x + P(e, flags[l], L.shape)
    for wk in reversed(w):
        x = x + wk * (b - L.apply(x)) / L.d
    return x.astype(np.float32).astype(float)

def cheb(lo, hi, k):
    return [1.0 / (0.5 * (hi + lo) + 0.5 * (hi - lo) * np.cos(np.pi * (2 * i + 1) / (2 * k))) for i in range(k)][::-1]

imgs = S.disc_images(6, 128)
configs = [("V(1,1) w.9 2x", 1, [0.9], False), ("V(2,2) cheb 2x", 1, [0.595, 1.614], False), ("V(2,2) cheb 2x half", 1, [0.595, 1.614], True),
           ("V(2,2) cheb[.25,1.9] 4x", 2, cheb(0.25, 1.9, 2), False), ("V(3,3) cheb[.15,1.9] 4x", 2, cheb(0.15, 1.9, 3), False),
           ("V(4,4) cheb[.1,1.9] 4x", 2, cheb(0.1, 1.9, 4), False), ("V(3,3) cheb[.25,1.9] 2x", 1, cheb(0.25, 1.9, 3), False)]
for name, jump, w, half in configs:
    its = []
    for img in imgs:
        solid = img[None]
        levels, flags, dense = build(solid, jump, half)
        L0 = levels[0]; nx = solid.shape[2]
        b = np.zeros(solid.shape); b[:, :, -1] = L0.wr
        x0 = np.broadcast_to((np.arange(nx) + 0.5) / nx, solid.shape).copy()
        x, it = pcg(L0, b, x0, lambda r: vcycle(levels, flags, 0, r, w, dense))
        its.append(it)
    nl = len(levels)
    sten = 2 * len(w) + 1   # fine-level stencil applications per iteration (first sweep is free: x=0) ~ 2nu - 1 + restrict + CG
    print(f"{name:28s} levels {[l.shape[1] for l in levels]} iters {its} mean {np.mean(its):.1f}  fine stencils/iter {sten}  total {np.mean(its)*sten:.0f}", flush=True)
