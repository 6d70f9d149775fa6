"""GPU tests of the multigrid-preconditioned grid solver: the preconditioner must be a symmetric
positive definite operator (CG's requirement), the multigrid path must land on the same discrete
solution as the plain diagonally preconditioned path and as the oracle, in far fewer iterations."""
import numpy as np
import pytest

from keff_cfd_b200 import structures as S

pytestmark = pytest.mark.gpu

SHAPES = [(1, 128, 128), (1, 34, 70), (24, 24, 24), (20, 33, 64), (7, 30, 50), (64, 64, 64), (3, 16, 200)]


@pytest.mark.parametrize("shape", SHAPES)
def test_vcycle_is_symmetric_positive_definite(K, shape):
    rng = np.random.default_rng(7)
    m = S.bernoulli_volume(shape, 0.5, 11)
    a, b = rng.standard_normal(shape), rng.standard_normal(shape)
    Ma, Mb = K.precondition(m, a), K.precondition(m, b)
    ab, ba = float((Ma * b).sum()), float((a * Mb).sum())
    # FP32 coarse levels: symmetric to single-precision rounding
    assert abs(ab - ba) <= 2e-5 * np.sqrt(float((Ma * a).sum()) * float((Mb * b).sum()))
    assert float((Ma * a).sum()) > 0 and float((Mb * b).sum()) > 0
    # linear: M(2a - 3b) = 2 Ma - 3 Mb
    Mc = K.precondition(m, 2 * a - 3 * b)
    assert np.abs(Mc - (2 * Ma - 3 * Mb)).max() <= 1e-5 * np.abs(Mc).max()


@pytest.mark.parametrize("shape", SHAPES)
def test_multigrid_path_matches_plain_cg_and_oracle(K, R, shape):
    m = S.bernoulli_volume(shape, 0.5, 5)
    pm = K.default_params(rel_residual=1e-10)
    pj = K.default_params(rel_residual=1e-10, flags=K.F_NO_MG)
    a, b = K.solve3d(m, pm), K.solve3d(m, pj)
    assert a["status"] == 0 and b["status"] == 0
    assert abs(a["keff"] - b["keff"]) <= 1e-7 * abs(b["keff"])
    assert np.abs(a["T"] - b["T"]).max() <= 1e-7
    assert a["rel_residual"] <= 1e-9
    assert a["iters"] < b["iters"]
    if np.prod(shape) <= 40000:
        d = R.orc_discrete(m if shape[0] > 1 else S.to_gray(m[0]))
        assert abs(a["keff"] - d["keff"]) <= 1e-7 * abs(d["keff"])
        assert np.abs(a["T"].ravel() - np.asarray(d["T"]).ravel()).max() <= 1e-6


def test_multigrid_iteration_count_is_grid_independent(K):
    its = []
    for n in (32, 64, 128):
        r = K.solve3d(S.bernoulli_volume((n, n, n), 0.5, 1234), K.default_params(), want_T=False)
        assert r["status"] == 0
        its.append(r["iters"])
    assert max(its) <= 40 and its[2] <= its[0] + 8, its


def test_multigrid_high_contrast_and_other_walls(K):
    m = S.bernoulli_volume((32, 32, 32), 0.4, 3)
    for kw in (dict(ks=100.0, kf=1.0), dict(ks=0.1, kf=3.0, tl=2.0, tr=-1.0)):
        a = K.solve3d(m, K.default_params(rel_residual=1e-10, **kw))
        b = K.solve3d(m, K.default_params(rel_residual=1e-10, flags=K.F_NO_MG, **kw))
        assert a["status"] == 0 and abs(a["keff"] - b["keff"]) <= 1e-8 * abs(b["keff"])
        assert a["iters"] < b["iters"]
