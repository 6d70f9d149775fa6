"""numpy restatement of the on-chip multigrid cycle of k_batch2d_mg (keff_cfd_b200/csrc/batch2d_mg.cuh): same
hierarchy (2x2 aggregates, halved coarse conductances, dense 8x8-cell level), same sweeps (one weighted-Jacobi sweep
per side on the image, two on the coarse levels), FP64 throughout.  Test infrastructure only: it documents the
algorithm the kernel implements and lets the CPU suite check that the cycle is symmetric positive definite and how
many CG iterations it needs; the kernel itself is checked against the oracle in tests/test_gpu_parity.py."""
import numpy as np


class Level:
    def __init__(self, cx, cy, wall):
        self.cx, self.cy, self.wall = cx, cy, wall          # faces towards x+1 (ny, nx-1) / y+1 (ny-1, nx); wall (ny, nx)
        d = wall.copy()
        d[:, :-1] += cx; d[:, 1:] += cx
        d[:-1, :] += cy; d[1:, :] += cy
        self.d = d

    def offdiag(self, x):                                   # sum_f c_f x_nb
        s = np.zeros_like(x)
        s[:, :-1] += self.cx * x[:, 1:]; s[:, 1:] += self.cx * x[:, :-1]
        s[:-1, :] += self.cy * x[1:, :]; s[1:, :] += self.cy * x[:-1, :]
        return s

    def apply(self, x):
        return self.d * x - self.offdiag(x)

    def coarsen(self):
        ny, nx = self.d.shape
        cx = 0.5 * (self.cx[0::2, 1::2] + self.cx[1::2, 1::2])[:, :nx // 2 - 1]
        cy = 0.5 * (self.cy[1::2, 0::2] + self.cy[1::2, 1::2])[:ny // 2 - 1, :]
        wall = 0.5 * (self.wall[0::2, 0::2] + self.wall[0::2, 1::2] + self.wall[1::2, 0::2] + self.wall[1::2, 1::2])
        return Level(cx, cy, wall)


def fine_level(solid, ks=10.0, kf=1.0):
    """The reference's 2D discretisation (Keff2D/keff2D.h:482-582) on the unit square."""
    ny, nx = solid.shape
    K = np.where(solid == 1, ks, kf).astype(float)
    dx, dy = 1.0 / nx, 1.0 / ny
    hm = lambda a, b: 2 * a * b / (a + b)
    cx = hm(K[:, :-1], K[:, 1:]) * dy / dx
    cy = hm(K[:-1, :], K[1:, :]) * dx / dy
    wall = np.zeros((ny, nx))
    wall[:, 0] += K[:, 0] * dy / (dx / 2)
    wall[:, -1] += K[:, -1] * dy / (dx / 2)
    return Level(cx, cy, wall)


def hierarchy(solid, ks=10.0, kf=1.0):
    levels = [fine_level(solid, ks, kf)]
    for _ in range(4):
        levels.append(levels[-1].coarsen())
    L = levels[-1]
    n = L.d.size
    A = np.zeros((n, n))
    for i in range(n):
        e = np.zeros(n); e[i] = 1
        A[:, i] = L.apply(e.reshape(L.d.shape)).ravel()
    return levels, np.linalg.inv(A)


def dense_matrix(L):
    """The dense-level matrix the kernel assembles in shared memory (5-point operator of the coarsest level)."""
    n = L.d.size
    A = np.zeros((n, n))
    for i in range(n):
        e = np.zeros(n); e[i] = 1
        A[:, i] = L.apply(e.reshape(L.d.shape)).ravel()
    return A


def block_gauss_jordan(A, dtype=np.float32, bs=8):
    """In-place inverse the way k_batch2d_mg forms it (batch2d_mg.cuh): the matrix padded to 64 x 64 with an identity,
    one step per 8 x 8 pivot block, no pivoting (SPD M-matrix).  Step kb with P = (pivot block)^-1, C = the pivot
    columns, R = the pivot rows:  pivot rows <- P R (pivot columns: P);  every other row i <- (row with its pivot
    columns zeroed) - C_i (P R with its pivot columns replaced by P).  P itself by eight scalar Gauss-Jordan steps."""
    na = A.shape[0]
    M = np.zeros((64, 64), dtype)
    M[:na, :na] = A
    for i in range(na, 64):
        M[i, i] = 1
    for kb in range((na + bs - 1) // bs):
        s = slice(bs * kb, bs * kb + bs)
        P = M[s, s].copy()
        for p in range(bs):
            piv = dtype(1) / P[p, p]
            row, col = P[p, :] * piv, P[:, p].copy()
            row[p] = piv
            keep = P.copy()
            keep[:, p] = 0
            P = (keep - np.outer(col, row)).astype(dtype)
            P[p, :] = row
        T = (P @ M[s, :]).astype(dtype)
        T[:, s] = P
        C = M[:, s].copy()
        keep = M.copy()
        keep[:, s] = 0
        new = (keep - C @ T).astype(dtype)
        new[s, :] = T
        M = new
    return M[:na, :na]


def restrict(r):
    return r[0::2, 0::2] + r[0::2, 1::2] + r[1::2, 0::2] + r[1::2, 1::2]


def prolong(e):
    return np.repeat(np.repeat(e, 2, 0), 2, 1)


def vcycle(levels, dense, b, l=0, om0=0.8, wc=(0.7, 1.4)):
    L = levels[l]
    if l == len(levels) - 1:
        return (dense @ b.ravel()).reshape(b.shape)
    w = (om0,) if l == 0 else wc
    x = np.zeros_like(b)
    for wk in w:
        x = x + wk * (b - L.apply(x)) / L.d
    x = x + prolong(vcycle(levels, dense, restrict(b - L.apply(x)), l + 1, om0, wc))
    for wk in reversed(w):
        x = x + wk * (b - L.apply(x)) / L.d
    return x


def pcg(solid, ks=10.0, kf=1.0, tol=1e-8, maxit=200, **kw):
    levels, dense = hierarchy(solid, ks, kf)
    L0 = levels[0]
    ny, nx = solid.shape
    b = np.zeros((ny, nx)); b[:, -1] = L0.wall[:, -1]          # TL = 0, TR = 1
    x = np.tile((np.arange(nx) + 0.5) / nx, (ny, 1))
    r = b - L0.apply(x); z = vcycle(levels, dense, r, **kw); p = z.copy(); rz = (r * z).sum(); bn = np.sqrt((b * b).sum())
    for it in range(1, maxit + 1):
        q = L0.apply(p); a = rz / (p * q).sum(); x += a * p; r -= a * q
        if np.sqrt((r * r).sum()) <= tol * bn:
            break
        z = vcycle(levels, dense, r, **kw); rzn = (r * z).sum(); p = z + (rzn / rz) * p; rz = rzn
    keff = 0.5 * ((L0.wall[:, 0] * x[:, 0]).sum() + (L0.wall[:, -1] * (1 - x[:, -1])).sum())
    return keff, it, x


def direct_keff(solid, ks=10.0, kf=1.0, tl=0.0, tr=1.0):
    """Keff from a sparse direct solve (scipy, FP64) of the same linear system: an answer that owes nothing to any of
    the iterative solvers or their stop rules."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as sla
    L = fine_level(solid, ks, kf)
    ny, nx = solid.shape
    idx = np.arange(nx * ny).reshape(ny, nx)
    rows = [idx.ravel(), idx[:, :-1].ravel(), idx[:, 1:].ravel(), idx[:-1, :].ravel(), idx[1:, :].ravel()]
    cols = [idx.ravel(), idx[:, 1:].ravel(), idx[:, :-1].ravel(), idx[1:, :].ravel(), idx[:-1, :].ravel()]
    vals = [L.d.ravel(), -L.cx.ravel(), -L.cx.ravel(), -L.cy.ravel(), -L.cy.ravel()]
    A = sp.csc_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(nx * ny, nx * ny))
    b = np.zeros((ny, nx))
    b[:, 0] += tl * L.wall[:, 0]
    b[:, -1] += tr * L.wall[:, -1]
    T = sla.spsolve(A, b.ravel()).reshape(ny, nx)
    ql = (L.wall[:, 0] * (T[:, 0] - tl)).sum()
    qr = (L.wall[:, -1] * (tr - T[:, -1])).sum()
    return (ql + qr) / 2 / (tr - tl)
