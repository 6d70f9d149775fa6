"""Prototype of the exact multigrid preconditioner the CUDA code implements: aggregates of 2 cells per
direction (ceil for odd sizes, directions of size 1 untouched), coarse face conductance = 0.5 x sum of the
fine faces crossing the coarse face (1.0 x for a direction that is not coarsened), one weighted-Jacobi
sweep before and after (V-cycle), dense solve at <= 64 cells.  Counts PCG iterations; numpy only."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keff_cfd_b200 import structures as S


def agg(a, axis, n):
    """sum pairs along axis (ceil: a lone last element stays)"""
    idx = np.arange(0, n, 2)
    return np.add.reduceat(a, idx, axis=axis)


class Level:
    def __init__(self, cx, cy, cz, wl, wr, shape):
        self.cx, self.cy, self.cz, self.wl, self.wr, self.shape = cx, cy, cz, wl, wr, shape
        d = np.zeros(shape)
        d[:, :, :-1] += cx; d[:, :, 1:] += cx
        d[:, :-1, :] += cy; d[:, 1:, :] += cy
        d[:-1] += cz; d[1:] += cz
        d[:, :, 0] += wl; d[:, :, -1] += wr
        self.d = d

    def apply(self, x):
        s = self.d * x
        s[:, :, :-1] -= self.cx * x[:, :, 1:]; s[:, :, 1:] -= self.cx * x[:, :, :-1]
        s[:, :-1, :] -= self.cy * x[:, 1:, :]; s[:, 1:, :] -= self.cy * x[:, :-1, :]
        s[:-1] -= self.cz * x[1:]; s[1:] -= self.cz * x[:-1]
        return s

    def coarsen(self):
        nz, ny, nx = self.shape
        f = [nz > 1, ny > 1, nx > 1]                       # coarsened directions (z, y, x)
        def reduce_all(a, skip_axis=None):
            for ax, n in enumerate(a.shape):
                if ax != skip_axis and f[ax]:
                    a = agg(a, ax, n)
            return a
        sc = lambda ax: 0.5 if f[ax] else 1.0
        # x faces crossing coarse boundaries: fine face index 1,3,5.. along x (between 2X+1 and 2X+2)
        ccx = sc(2) * reduce_all(self.cx[:, :, 1::2], 2) if f[2] else self.cx.copy()
        ccy = sc(1) * reduce_all(self.cy[:, 1::2, :], 1) if f[1] else reduce_all(self.cy, 1)
        ccz = sc(0) * reduce_all(self.cz[1::2], 0) if f[0] else reduce_all(self.cz, 0)
        def r2(a):  # wall arrays (nz, ny)
            if f[0]: a = agg(a, 0, a.shape[0])
            if f[1]: a = agg(a, 1, a.shape[1])
            return a
        wl, wr = sc(2) * r2(self.wl), sc(2) * r2(self.wr)
        shape = tuple((n + 1) // 2 if ff else n for n, ff in zip(self.shape, f))
        # trim face arrays to (n-1) along their own axis
        ccx = ccx[:, :, :shape[2] - 1]; ccy = ccy[:, :shape[1] - 1, :]; ccz = ccz[:shape[0] - 1]
        return Level(ccx, ccy, ccz, wl, wr, shape), f


def restrict(r, f):
    for ax in range(3):
        if f[ax]: r = agg(r, ax, r.shape[ax])
    return r


def prolong(e, f, shape):
    for ax in range(3):
        if f[ax]: e = np.repeat(e, 2, ax).take(range(shape[ax]), axis=ax)
    return e


def fine_level(solid, ks=10.0, kf=1.0):
    nz, ny, nx = solid.shape
    K = np.where(solid == 1, ks, kf).astype(float)
    hm = lambda a, b: 2 * a * b / (a + b)
    dx, dy, dz = 1.0 / nx, 1.0 / ny, (1.0 / nz if nz > 1 else 1.0)
    cx = hm(K[:, :, :-1], K[:, :, 1:]) * dy * dz / dx
    cy = hm(K[:, :-1, :], K[:, 1:, :]) * dx * dz / dy
    cz = hm(K[:-1], K[1:]) * dx * dy / dz
    wl = K[:, :, 0] * dy * dz / (dx / 2); wr = K[:, :, -1] * dy * dz / (dx / 2)
    return Level(cx, cy, cz, wl, wr, solid.shape)


def dense_inverse(L):
    n = int(np.prod(L.shape)); A = np.zeros((n, n))
    for i in range(n):
        e = np.zeros(n); e[i] = 1; A[:, i] = L.apply(e.reshape(L.shape)).ravel()
    return np.linalg.inv(A)


def vcycle(levels, flags, l, b, omega, dense, f32):
    L = levels[l]
    if l == len(levels) - 1:
        return (dense @ b.ravel()).reshape(b.shape)
    dt = np.float32 if (f32 and l > 0) else np.float64
    u = (omega * b / L.d).astype(dt).astype(float)
    res = b - L.apply(u)
    bc = restrict(res, flags[l]).astype(np.float32 if f32 else float).astype(float)
    e = vcycle(levels, flags, l + 1, bc, omega, dense, f32)
    v = u + prolong(e, flags[l], L.shape)
    return (v + omega * (b - L.apply(v)) / L.d).astype(dt).astype(float)


def pcg(L, b, x0, M, tol=1e-8, maxit=500):
    x = x0.copy(); r = b - L.apply(x); z = M(r); p = z.copy(); rz = (r * z).sum(); bn = np.sqrt((b * b).sum())
    for it in range(1, maxit + 1):
        q = L.apply(p); a = rz / (p * q).sum(); x += a * p; r -= a * q
        if np.sqrt((r * r).sum()) <= tol * bn: return x, it
        z = M(r); rzn = (r * z).sum(); p = z + (rzn / rz) * p; rz = rzn
    return x, maxit


def run(solid, name, omega=0.9, ks=10.0, f32=True):
    L0 = fine_level(solid, ks=ks)
    nz, ny, nx = solid.shape
    b = np.zeros(solid.shape); b[:, :, -1] = L0.wr
    x0 = np.broadcast_to((np.arange(nx) + 0.5) / nx, solid.shape).copy()
    levels, flags = [L0], []
    while np.prod(levels[-1].shape) > 64:
        Lc, f = levels[-1].coarsen(); levels.append(Lc); flags.append(f)
    dense = dense_inverse(levels[-1])
    x, it = pcg(L0, b, x0, lambda r: vcycle(levels, flags, 0, r, omega, dense, f32))
    _, itj = pcg(L0, b, x0, lambda r: r / L0.d, maxit=5000)
    keff = 0.5 * ((L0.wl * x[:, :, 0]).sum() + (L0.wr * (1 - x[:, :, -1])).sum())
    print(name, solid.shape, [l.shape for l in levels[-2:]], "omega", omega, "mg iters", it, "jacobi iters", itj, "keff", round(keff, 9), flush=True)


if __name__ == "__main__":
    for shp in ((32, 32, 32), (1, 128, 128), (1, 100, 60), (7, 30, 50), (20, 33, 64), (64, 64, 64), (1, 256, 256)):
        run(S.bernoulli_volume(shp, 0.5, 1234), "bern")
    run(S.disc_images(1, 128)[0][None], "disc128")
    for om in (0.8, 1.0):
        run(S.bernoulli_volume((32, 32, 32), 0.5, 1234), "bern", omega=om)
