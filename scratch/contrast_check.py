"""Who is right at high contrast?  Keff of the on-chip multigrid kernel and of the all-FP64 Jacobi-PCG kernel against a
sparse direct solve of the same system (scipy, FP64) on 96x112 random structures, ks:kf = 300."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "helpers"))
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as sla
import onchip_mg_proto as P

def direct(img, ks, kf, tl, tr):
    L = P.fine_level(img, ks, kf)
    ny, nx = img.shape; n = nx * ny
    idx = np.arange(n).reshape(ny, nx)
    rows = [idx.ravel()]; cols = [idx.ravel()]; vals = [L.d.ravel()]
    rows += [idx[:, :-1].ravel(), idx[:, 1:].ravel()]; cols += [idx[:, 1:].ravel(), idx[:, :-1].ravel()]; vals += [-L.cx.ravel(), -L.cx.ravel()]
    rows += [idx[:-1, :].ravel(), idx[1:, :].ravel()]; cols += [idx[1:, :].ravel(), idx[:-1, :].ravel()]; vals += [-L.cy.ravel(), -L.cy.ravel()]
    A = sp.csc_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n))
    wl = np.zeros((ny, nx)); wl[:, 0] = L.wall[:, 0]; wr = np.zeros((ny, nx)); wr[:, -1] = L.wall[:, -1]
    b = (tl * wl + tr * wr).ravel()
    T = sla.spsolve(A, b).reshape(ny, nx)
    ql = (L.wall[:, 0] * (T[:, 0] - tl)).sum(); qr = (L.wall[:, -1] * (tr - T[:, -1])).sum()
    return (ql + qr) / 2 / (tr - tl)

from keff_cfd_b200 import structures as S
rng = np.random.default_rng(3)
discs = S.disc_images(8, 128)
target = (96, 112) if len(sys.argv) < 3 else (int(sys.argv[2]), int(sys.argv[3]))
for ny in range(16, 129, 16):              # the image sequence of scratch/stress_batch.py
    for nx in range(16, 129, 16):
        cand = np.concatenate([(rng.random((3, ny, nx)) < p).astype(np.uint8) for p in (0.2, 0.5, 0.8)] + [discs[:3, :ny, :nx]])
        if (ny, nx) == target:
            imgs = cand
ny, nx = target
ks, kf, tl, tr = 300.0, 1.0, 1.0, 3.0
ref = np.array([direct(im, ks, kf, tl, tr) for im in imgs])
if len(sys.argv) > 1 and sys.argv[1] == "gpu":
    import keff_cfd_b200 as K
    K.init(0)
    p = dict(ks=ks, kf=kf, tl=tl, tr=tr)
    a = K.solve2d_batch(imgs, K.default_params(**p)); b = K.solve2d_batch(imgs, K.default_params(flags=K.F_BATCH_GLOBAL, **p))
    print("mg   gap vs direct", np.abs(a["keff"] - ref) / ref, a["iters"])
    print("fp64 gap vs direct", np.abs(b["keff"] - ref) / ref, b["iters"])
    for tolr in (1e-10,):
        a = K.solve2d_batch(imgs, K.default_params(rel_residual=tolr, **p)); b = K.solve2d_batch(imgs, K.default_params(flags=K.F_BATCH_GLOBAL, rel_residual=tolr, **p))
        print(f"rel_residual {tolr}: mg", np.abs(a["keff"] - ref).max() / ref.max(), "fp64", (np.abs(b["keff"] - ref) / ref).max())
else:
    print(ref)
if len(sys.argv) > 1 and sys.argv[1] == "gpu":
    g = np.array([K.solve2d(im, K.default_params(**p), want_T=False)["keff"] for im in imgs])
    gj = np.array([K.solve2d(im, K.default_params(flags=K.F_NO_MG, **p), want_T=False)["keff"] for im in imgs])
    print("grid solver (multigrid-PCG) gap vs direct", (np.abs(g - ref) / ref).max(), "| Jacobi-PCG fallback", (np.abs(gj - ref) / ref).max())
    vol = np.stack([imgs[0], imgs[3], imgs[6], imgs[9]])          # a 4-plane 3D volume, direct solve too big: compare the two paths
    a3 = K.solve3d(vol, K.default_params(**p), want_T=False); b3 = K.solve3d(vol, K.default_params(flags=K.F_NO_MG, **p), want_T=False)
    print("3D 4x96x112: mg", a3["keff"], a3["iters"], "jacobi", b3["keff"], b3["iters"], "gap", abs(a3["keff"] - b3["keff"]) / b3["keff"])
