"""Prototype: geometric multigrid (cell-centred, piecewise-constant transfer, scaled Galerkin coarse
operators) as a CG preconditioner for the 3D conduction system.  numpy only; counts iterations."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keff_cfd_b200 import structures as S


class Level:
    def __init__(self, cx, cy, cz, wl, wr):
        self.cx, self.cy, self.cz, self.wl, self.wr = cx, cy, cz, wl, wr
        nz, ny, nx = cy.shape[0], cx.shape[1], cz.shape[2]
        self.shape = (nz, ny, nx)
        d = np.zeros(self.shape)
        d[:, :, :-1] += cx; d[:, :, 1:] += cx
        d[:, :-1, :] += cy; d[:, 1:, :] += cy
        d[:-1] += cz; d[1:] += cz
        d[:, :, 0] += wl; d[:, :, -1] += wr
        self.d = d
        zz, yy, xx = np.indices(self.shape)
        self.red = ((zz + yy + xx) % 2 == 0)

    def offdiag(self, x):
        s = np.zeros_like(x)
        s[:, :, :-1] += self.cx * x[:, :, 1:]; s[:, :, 1:] += self.cx * x[:, :, :-1]
        s[:, :-1, :] += self.cy * x[:, 1:, :]; s[:, 1:, :] += self.cy * x[:, :-1, :]
        s[:-1] += self.cz * x[1:]; s[1:] += self.cz * x[:-1]
        return s

    def apply(self, x):
        return self.d * x - self.offdiag(x)

    def coarsen(self, scale):
        cx, cy, cz = self.cx, self.cy, self.cz
        def blk(a):  # sum 2x2x2 blocks over (z,y,x) of an array with even dims
            return a.reshape(a.shape[0] // 2, 2, a.shape[1] // 2, 2, a.shape[2] // 2, 2)
        # faces crossing coarse boundaries: fine x-face index 1,3,5,... (between 2K+1 and 2K+2)
        fx = cx[:, :, 1::2]                                   # (nz, ny, nx/2-1)
        ccx = scale * fx.reshape(fx.shape[0] // 2, 2, fx.shape[1] // 2, 2, fx.shape[2]).sum(axis=(1, 3))
        fy = cy[:, 1::2, :]
        ccy = scale * fy.reshape(fy.shape[0] // 2, 2, fy.shape[1], fy.shape[2] // 2, 2).sum(axis=(1, 4))
        fz = cz[1::2]
        ccz = scale * fz.reshape(fz.shape[0], fz.shape[1] // 2, 2, fz.shape[2] // 2, 2).sum(axis=(2, 4))
        wl = scale * self.wl.reshape(self.wl.shape[0] // 2, 2, self.wl.shape[1] // 2, 2).sum(axis=(1, 3))
        wr = scale * self.wr.reshape(self.wr.shape[0] // 2, 2, self.wr.shape[1] // 2, 2).sum(axis=(1, 3))
        return Level(ccx, ccy, ccz, wl, wr)


def restrict(r):
    return r.reshape(r.shape[0] // 2, 2, r.shape[1] // 2, 2, r.shape[2] // 2, 2).sum(axis=(1, 3, 5))


def prolong(e):
    return np.repeat(np.repeat(np.repeat(e, 2, 0), 2, 1), 2, 2)


def fine_level(solid, ks=10.0, kf=1.0):
    nz, ny, nx = solid.shape
    K = np.where(solid == 1, ks, kf).astype(float)
    hm = lambda a, b: 2 * a * b / (a + b)
    dx, dy, dz = 1.0 / nx, 1.0 / ny, 1.0 / nz
    cx = hm(K[:, :, :-1], K[:, :, 1:]) * dy * dz / dx
    cy = hm(K[:, :-1, :], K[:, 1:, :]) * dx * dz / dy
    cz = hm(K[:-1], K[1:]) * dx * dy / dz
    wl = K[:, :, 0] * dy * dz / (dx / 2); wr = K[:, :, -1] * dy * dz / (dx / 2)
    return Level(cx, cy, cz, wl, wr)


def smooth(L, x, b, kind, nu, forward, omega=0.8):
    for _ in range(nu):
        if kind == "jacobi":
            x = x + omega * (b - L.apply(x)) / L.d
        else:  # red-black Gauss-Seidel; forward = red then black, backward = black then red
            for colour in ((True, False) if forward else (False, True)):
                m = L.red if colour else ~L.red
                x = np.where(m, (b + L.offdiag(x)) / L.d, x)
    return x


def vcycle(levels, l, b, kind, nu, dense):
    L = levels[l]
    if l == len(levels) - 1:
        return (dense @ b.ravel()).reshape(b.shape)
    x = smooth(L, np.zeros_like(b), b, kind, nu, True)
    r = b - L.apply(x)
    e = vcycle(levels, l + 1, restrict(r), kind, nu, dense)
    x = x + prolong(e)
    return smooth(L, x, b, kind, nu, False)


def dense_inverse(L):
    n = np.prod(L.shape)
    A = np.zeros((n, n))
    for i in range(n):
        e = np.zeros(n); e[i] = 1
        A[:, i] = L.apply(e.reshape(L.shape)).ravel()
    return np.linalg.inv(A)


def pcg(L, b, x0, M, tol=1e-8, maxit=3000):
    x = x0.copy(); r = b - L.apply(x); z = M(r); p = z.copy(); rz = (r * z).sum(); bn = np.sqrt((b * b).sum())
    for it in range(1, maxit + 1):
        q = L.apply(p); a = rz / (p * q).sum(); x += a * p; r -= a * q
        if np.sqrt((r * r).sum()) <= tol * bn: return x, it
        z = M(r); rzn = (r * z).sum(); p = z + (rzn / rz) * p; rz = rzn
    return x, maxit


def run(solid, name):
    L0 = fine_level(solid)
    nz, ny, nx = solid.shape
    b = np.zeros(solid.shape); b[:, :, -1] = L0.wr * 1.0   # TL = 0, TR = 1
    x0 = np.broadcast_to((np.arange(nx) + 0.5) / nx, solid.shape).copy()
    out = {}
    t = time.time(); _, out["jacobi"] = pcg(L0, b, x0, lambda r: r / L0.d); tj = time.time() - t
    for scale in (0.5, 1.0):
        levels = [L0]
        while min(levels[-1].shape) > 4 and all(s % 2 == 0 for s in levels[-1].shape):
            levels.append(levels[-1].coarsen(scale))
        dense = dense_inverse(levels[-1])
        for kind, nu in (("rbgs", 1), ("rbgs", 2), ("jacobi", 2)):
            t = time.time()
            x, it = pcg(L0, b, x0, lambda r: vcycle(levels, 0, r, kind, nu, dense))
            keff = 0.5 * ((L0.wl * x[:, :, 0]).sum() + (L0.wr * (1 - x[:, :, -1])).sum())
            out[f"mg s{scale} {kind}{nu}"] = (it, round(time.time() - t, 1), round(keff, 8))
    print(name, solid.shape, "levels", len(levels), out, flush=True)


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    run(S.bernoulli_volume((n, n, n), 0.5, 1234), f"bernoulli{n}")
    if n >= 64:
        m = np.unpackbits(np.load("tests/golden/schwarz163_bits.npy")).reshape(128, 128, 128)
        run(m[:n, :n, :n].copy() if n < 128 else m, "schwarz163")
