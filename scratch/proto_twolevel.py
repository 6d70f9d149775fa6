import sys, os, numpy as np, scipy.sparse as sp, scipy.sparse.linalg as sla
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keff_cfd_b200 import structures as S
from oracle import refbind as R

def system(img):
    A5, b, K = R.orc_assemble2d(S.to_gray(img))
    ny, nx = img.shape; n = nx*ny
    off = [0, -1, 1, nx, -nx]
    rows, cols, vals = [], [], []
    idx = np.arange(n)
    for j in range(5):
        m = A5[:, j] != 0
        rows.append(idx[m]); cols.append(idx[m] + off[j]); vals.append(A5[m, j])
    A = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n))
    return A, b

def pcg(A, b, Minv, tol=1e-8, x0=None, maxit=5000):
    x = np.zeros_like(b) if x0 is None else x0.copy()
    r = b - A @ x; z = Minv(r); p = z.copy(); rz = r @ z; bn = np.linalg.norm(b)
    for it in range(1, maxit+1):
        q = A @ p; a = rz / (p @ q); x += a*p; r -= a*q
        if np.linalg.norm(r) <= tol*bn: return x, it
        z = Minv(r); rzn = r @ z; p = z + (rzn/rz)*p; rz = rzn
    return x, maxit

for i in range(4):
    img = S.disc_image(i); ny, nx = img.shape
    A, b = system(img); d = A.diagonal()
    x0 = np.tile((np.arange(nx)+0.5)/nx, ny)
    res = {}
    _, res['jacobi'] = pcg(A, b, lambda r: r/d, x0=x0)
    for ag in (32, 16, 8):
        gy, gx = np.mgrid[0:ny, 0:nx]; agg = ((gy//ag)*(nx//ag) + gx//ag).ravel(); na = agg.max()+1
        P = sp.csr_matrix((np.ones(nx*ny), (np.arange(nx*ny), agg)), shape=(nx*ny, na))
        Ac = (P.T @ A @ P).toarray(); Aci = np.linalg.inv(Ac)
        _, res[f'add{ag}'] = pcg(A, b, lambda r: r/d + P @ (Aci @ (P.T @ r)), x0=x0)
        # multiplicative variant (coarse first then jacobi on updated residual, symmetrised)
        def mult(r):
            e = P @ (Aci @ (P.T @ r)); z = e + (r - A @ e)/d; return z + P @ (Aci @ (P.T @ (r - A @ z)))
        _, res[f'mul{ag}'] = pcg(A, b, mult, x0=x0)
    print(i, round(img.mean(),2), res)
