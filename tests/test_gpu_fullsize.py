"""BASELINE.json's full sizes on the GPU, checked through size-independent properties (no CPU oracle can solve
these in test time): convergence records, wall-flux balance, the Wiener bounds every two-phase structure obeys
(harmonic mean <= Keff <= arithmetic mean of the cell conductivities), mirror symmetries, independence of the wall
temperatures (the problem is linear), and agreement of the two independent solver paths on a sample."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

KS, KF = 10.0, 1.0


def wiener(phi):
    return 1.0 / (phi / KS + (1.0 - phi) / KF), phi * KS + (1.0 - phi) * KF


def test_batch_of_10000_images_properties(K):
    """configs[1]: 10 000 synthetic 128x128 images (the bench workload, same generator and seeds)."""
    from keff_cfd_b200 import structures as S
    imgs = S.disc_images(10000, 128)
    res = K.solve2d_batch(imgs)
    assert (res["status"] == 0).all() and (res["rel_residual"] <= 1e-8).all()
    assert (np.abs(res["q_left"] - res["q_right"]) <= 1e-6 * res["keff"]).all()          # energy conservation
    phi = imgs.reshape(len(imgs), -1).mean(axis=1)
    lo, hi = wiener(phi)
    assert (res["keff"] >= lo * (1 - 1e-9)).all() and (res["keff"] <= hi * (1 + 1e-9)).all()
    assert res["iters"].max() <= 40 and res["iters"].mean() < 16                          # multigrid: size-independent count
    sub = slice(0, 10000, 20)                                                              # 500 images
    base = res["keff"][sub]
    flip_y = K.solve2d_batch(np.ascontiguousarray(imgs[sub][:, ::-1, :]))["keff"]          # adiabatic faces swapped
    flip_x = K.solve2d_batch(np.ascontiguousarray(imgs[sub][:, :, ::-1]))["keff"]          # walls swapped
    other_t = K.solve2d_batch(imgs[sub], K.default_params(tl=350.0, tr=290.0))["keff"]     # linear in (TL, TR)
    fp64 = K.solve2d_batch(imgs[sub][:50], K.default_params(flags=K.F_BATCH_GLOBAL))["keff"]  # other solver path
    for other in (flip_y, flip_x, other_t):
        assert (np.abs(other - base) <= 5e-6 * base).all()   # stop rule: residual 1e-8 of ||b||, wall fluxes equal to 1e-6
    assert (np.abs(fp64 - base[:50]) <= 5e-6 * fp64).all()


def test_volume_512_cubed_properties(K):
    """configs[3]: random binary 512^3 volume, ks:kf = 10:1, one GPU (134 M cells, 7.4 GB of workspace)."""
    import torch
    from keff_cfd_b200 import api
    vol = api.bernoulli_volume_dev((512, 512, 512), 0.5, 1234)
    r = api.solve3d(vol, want_T=True)
    assert r["status"] == 0 and r["rel_residual"] <= 1e-8 and r["iters"] <= 12
    assert abs(r["q_left"] - r["q_right"]) <= 1e-6 * r["keff"]
    lo, hi = wiener(float(vol.float().mean()))
    assert lo < r["keff"] < hi
    # same medium statistics as the numpy-seeded goldens of SURVEY 8c (32^3: 3.0440, 64^3: 3.0197, 128^3: 3.0153)
    assert abs(r["keff"] - 3.0146) < 5e-3
    T = r["T"]
    assert float(T.min()) >= -1e-9 and float(T.max()) <= 1 + 1e-9                         # discrete maximum principle
    # the wall fluxes recomputed from the returned field by torch: Keff = (QL + QR) / 2 / (TR - TL)
    kcell = torch.where(vol == 1, KS, KF).double()
    cw = 2.0 / 512                                   # k dy dz / (dx / 2) with dx = dy = dz = 1/512, per unit k
    ql = float((kcell[:, :, 0] * cw * T[:, :, 0]).sum())
    qr = float((kcell[:, :, -1] * cw * (1.0 - T[:, :, -1])).sum())
    assert abs((ql + qr) / 2 - r["keff"]) <= 1e-9 * r["keff"]
    del T, kcell, r
    mirrored = torch.flip(vol, dims=[2]).contiguous()                                     # walls swapped
    rm = api.solve3d(mirrored, want_T=False)
    assert rm["status"] == 0 and abs(rm["keff"] - (ql + qr) / 2) <= 1e-6 * rm["keff"]
    transposed = vol.permute(1, 0, 2).contiguous()                                        # y <-> z: same conductance along x
    rt = api.solve3d(transposed, want_T=False)
    assert rt["status"] == 0 and abs(rt["keff"] - (ql + qr) / 2) <= 1e-6 * rt["keff"]
