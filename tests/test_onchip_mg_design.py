"""CPU checks of the algorithm k_batch2d_mg implements (numpy restatement in tests/helpers/onchip_mg_proto.py):
the cycle is a symmetric positive definite operator, CG preconditioned with it lands on the oracle's discrete
solution, and it needs an order of magnitude fewer iterations than the diagonally preconditioned iteration.  (The
kernel itself is checked on the GPU: tests/test_gpu_parity.py.)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "helpers"))
import onchip_mg_proto as P  # noqa: E402
from keff_cfd_b200 import structures as S  # noqa: E402


def test_cycle_is_symmetric_positive_definite_and_linear():
    img = S.disc_image(3)
    levels, dense = P.hierarchy(img)
    assert [lv.d.shape for lv in levels] == [(128, 128), (64, 64), (32, 32), (16, 16), (8, 8)]
    rng = np.random.default_rng(1)
    a, b = rng.standard_normal(img.shape), rng.standard_normal(img.shape)
    Ma, Mb = P.vcycle(levels, dense, a), P.vcycle(levels, dense, b)
    assert abs((Ma * b).sum() - (a * Mb).sum()) <= 1e-12 * np.sqrt((Ma * a).sum() * (Mb * b).sum())
    assert (Ma * a).sum() > 0 and (Mb * b).sum() > 0
    assert np.abs(P.vcycle(levels, dense, 2 * a - 3 * b) - (2 * Ma - 3 * Mb)).max() <= 1e-12 * np.abs(Ma).max()


def test_cycle_preconditioned_cg_reaches_the_oracle_solution(R):
    for i, (ny, nx, cap) in enumerate([(128, 128, 16), (64, 96, 22)]):
        img = S.disc_image(i)[:ny, :nx]
        keff, it, T = P.pcg(img, tol=1e-9)
        d = R.orc_discrete(S.to_gray(img))
        assert abs(keff - d["keff"]) / d["keff"] < 1e-7 and np.abs(T - d["T"]).max() < 1e-6
        assert it <= cap      # relative residual 1e-9; the GPU kernel reports 12-14 (square cells) at 1e-8


def test_two_sweeps_on_the_coarse_levels_pay():
    img = S.disc_image(0)
    _, it22, _ = P.pcg(img)
    _, it11, _ = P.pcg(img, wc=(0.8,))                      # one sweep per side everywhere
    assert it22 < it11 <= 22


def test_block_gauss_jordan_is_the_dense_level_inverse():
    """The kernel inverts the coarsest-level matrix by block Gauss-Jordan with 8 x 8 pivot blocks in FP32, padded with
    an identity: same inverse as LAPACK's to single precision, symmetric, for every dense-level shape it meets."""
    for ny, nx in ((128, 128), (48, 80), (16, 16), (128, 16), (112, 128)):
        img = S.disc_image(4)[:ny, :nx]
        levels, _ = P.hierarchy(img)
        A = P.dense_matrix(levels[-1])
        inv = P.block_gauss_jordan(A)
        ref = np.linalg.inv(A)
        assert inv.shape == (ny // 16 * (nx // 16),) * 2
        assert np.abs(inv - ref).max() <= 2e-6 * np.abs(ref).max()
        assert np.abs(inv - inv.T).max() <= 2e-6 * np.abs(ref).max()
