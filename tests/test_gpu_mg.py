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


def test_fp32_search_direction_lands_on_the_same_solution(K):
    """The multigrid path stores p in FP32 (x and r advance with the same stored p, so r stays the residual of x);
    KEFFB200_P64=1 keeps it in FP64.  Same iteration count, same field to solver tolerance -- and the stencil bench
    entry point accepts the FP32-input form the iteration launches."""
    import json
    import os
    import subprocess
    import sys
    from conftest import ROOT
    prog = ("import json, sys, numpy as np; sys.path.insert(0, %r); import keff_cfd_b200 as K; "
            "from keff_cfd_b200 import structures as S; K.init(0); "
            "m = S.bernoulli_volume((40, 48, 64), 0.5, 21); r = K.solve3d(m, K.default_params(rel_residual=1e-11)); "
            "print(json.dumps({'keff': r['keff'], 'iters': r['iters'], 'status': r['status'], "
            "'T': [float(r['T'].min()), float(r['T'].max()), float(np.abs(r['T']).sum())]}))" % ROOT)
    out = {}
    for tag, env in (("p32", {}), ("p64", {"KEFFB200_P64": "1"})):
        r = subprocess.run([sys.executable, "-c", prog], capture_output=True, text=True, env={**os.environ, **env}, timeout=300)
        assert r.returncode == 0, r.stderr[-500:]
        out[tag] = json.loads(r.stdout.strip().splitlines()[-1])
    a, b = out["p32"], out["p64"]
    assert a["status"] == 0 and b["status"] == 0 and abs(a["iters"] - b["iters"]) <= 1
    assert abs(a["keff"] - b["keff"]) <= 1e-10 * abs(b["keff"])
    assert abs(a["T"][2] - b["T"][2]) <= 1e-9 * b["T"][2]
    import torch
    from keff_cfd_b200 import api
    vol = api.bernoulli_volume_dev((64, 64, 128), 0.5, 5)
    assert api.bench_apply(vol, api.default_params(flags=32), reps=3) > 0
    with pytest.raises(Exception):
        api.bench_apply(api.bernoulli_volume_dev((16, 16, 18), 0.5, 5), api.default_params(flags=32), reps=1)   # nx % 4 != 0
