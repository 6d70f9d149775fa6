"""TEST INFRASTRUCTURE ONLY: ctypes bindings for the checkers under oracle/.

  * ``ref2d`` / ``ref3d``   - the UNMODIFIED reference CPU code compiled from /root/reference by
                              oracle/Makefile into oracle/_ref/libkeffref{2d,3d}.so (shims:
                              ref_wrap2d.cpp / ref_wrap3d.cpp).
  * ``oracle``              - this repo's plain-C restatement (oracle/keff_oracle.c -> liboracle.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product package (keff_cfd_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_dp = C.POINTER(C.c_double)
_u8p = C.POINTER(C.c_ubyte)


def _p(a, typ):
    return None if a is None else a.ctypes.data_as(typ)


def build(quiet=True):
    """(Re)build liboracle.so and, when /root/reference is present, oracle/_ref."""
    subprocess.run(["make", "-C", HERE, "all"], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def have_ref():
    return all(os.path.exists(os.path.join(HERE, "_ref", f))
               for f in ("libkeffref2d.so", "libkeffref3d.so"))


_cache = {}


def _load(path):
    if path not in _cache:
        _cache[path] = C.CDLL(path)
    return _cache[path]


# --------------------------------------------------------------------------- reference (unmodified)
def ref2d_solve(img, amp=(1, 1), ks=10.0, kf=1.0, TL=0.0, TR=1.0, tol=1e-5, max_iter=500000,
                solver="serial", threads=1, want_T=True):
    """Run the reference 2D path (keff2D.cpp:102-170) on a u8 image (pixel<150 = fluid)."""
    lib = _load(os.path.join(HERE, "_ref", "libkeffref2d.so"))
    img = np.ascontiguousarray(img, dtype=np.uint8)
    H, W = img.shape
    nx, ny = W * amp[0], H * amp[1]
    T = np.zeros((ny, nx)) if want_T else None
    qL, qR, out = np.zeros(ny), np.zeros(ny), np.zeros(6)
    lib.ref2d_solve.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double,
                                C.c_double, C.c_double, C.c_double, C.c_long, C.c_int, C.c_int,
                                _dp, _dp, _dp, _dp]
    lib.ref2d_solve(_p(img, _u8p), W, H, amp[0], amp[1], ks, kf, TL, TR, tol, max_iter,
                    1 if solver == "parallel" else 0, threads, _p(T, _dp), _p(qL, _dp), _p(qR, _dp),
                    _p(out, _dp))
    return dict(keff=out[0], Q1=out[1], Q2=out[2], iters=int(out[3]), residual=out[4],
                porosity=out[5], T=T, qL=qL, qR=qR)


def ref2d_discretize(img, amp=(1, 1), ks=10.0, kf=1.0, TL=0.0, TR=1.0):
    lib = _load(os.path.join(HERE, "_ref", "libkeffref2d.so"))
    img = np.ascontiguousarray(img, dtype=np.uint8)
    H, W = img.shape
    n = W * amp[0] * H * amp[1]
    A, b, K = np.zeros((n, 5)), np.zeros(n), np.zeros(n)
    lib.ref2d_discretize.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double,
                                     C.c_double, C.c_double, _dp, _dp, _dp]
    lib.ref2d_discretize(_p(img, _u8p), W, H, amp[0], amp[1], ks, kf, TL, TR, _p(A, _dp), _p(b, _dp),
                         _p(K, _dp))
    return A, b, K


def ref2d_batch(imgs, ks=10.0, kf=1.0, TL=0.0, TR=1.0, tol=1e-5, max_iter=500000, threads=1):
    """Reference batch mode (keff2D.cpp:341-471): OpenMP over images, SerialGS each."""
    lib = _load(os.path.join(HERE, "_ref", "libkeffref2d.so"))
    imgs = np.ascontiguousarray(imgs, dtype=np.uint8)
    n, H, W = imgs.shape
    res = np.zeros((n, 4))
    lib.ref2d_batch.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double,
                                C.c_double, C.c_double, C.c_long, C.c_int, _dp]
    lib.ref2d_batch(_p(imgs, _u8p), n, W, H, ks, kf, TL, TR, tol, max_iter, threads, _p(res, _dp))
    return res


def ref3d_solve(solid, ks=10.0, kf=1.0, TL=0.0, TR=1.0, tol=1e-5, max_iter=500000, threads=1,
                want_T=True):
    """Run the reference 3D path (keff3D.cpp:99-163) on a u8 [nz,ny,nx] array (1 = solid)."""
    lib = _load(os.path.join(HERE, "_ref", "libkeffref3d.so"))
    solid = np.ascontiguousarray(solid, dtype=np.uint8)
    nz, ny, nx = solid.shape
    T = np.zeros((nz, ny, nx)) if want_T else None
    out = np.zeros(4)
    lib.ref3d_solve.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double,
                                C.c_double, C.c_double, C.c_long, C.c_int, _dp, _dp]
    lib.ref3d_solve(_p(solid, _u8p), nx, ny, nz, ks, kf, TL, TR, tol, max_iter, threads, _p(T, _dp),
                    _p(out, _dp))
    return dict(keff=out[0], Q1=out[1], Q2=out[2], iters=int(out[3]), T=T)


def ref3d_discretize(solid, ks=10.0, kf=1.0, TL=0.0, TR=1.0):
    lib = _load(os.path.join(HERE, "_ref", "libkeffref3d.so"))
    solid = np.ascontiguousarray(solid, dtype=np.uint8)
    nz, ny, nx = solid.shape
    n = nx * ny * nz
    A, b = np.zeros((n, 7)), np.zeros(n)
    lib.ref3d_discretize.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double,
                                     C.c_double, C.c_double, _dp, _dp]
    lib.ref3d_discretize(_p(solid, _u8p), nx, ny, nz, ks, kf, TL, TR, _p(A, _dp), _p(b, _dp))
    return A, b


# --------------------------------------------------------------------------- this repo's restatement
def _orc():
    path = os.path.join(HERE, "liboracle.so")
    if not os.path.exists(path):
        subprocess.run(["make", "-C", HERE, "liboracle.so"], check=True, stdout=subprocess.DEVNULL)
    return _load(path)


def orc_assemble2d(img, amp=(1, 1), ks=10.0, kf=1.0, TL=0.0, TR=1.0):
    lib = _orc()
    img = np.ascontiguousarray(img, dtype=np.uint8)
    H, W = img.shape
    nx, ny = W * amp[0], H * amp[1]
    n = nx * ny
    K, A, b = np.zeros(n), np.zeros((n, 5)), np.zeros(n)
    lib.orc_kmap2d.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, _dp]
    lib.orc_kmap2d(_p(img, _u8p), W, H, amp[0], amp[1], ks, kf, _p(K, _dp))
    lib.orc_assemble2d.argtypes = [_dp, C.c_int, C.c_int, C.c_double, C.c_double, _dp, _dp]
    lib.orc_assemble2d(_p(K, _dp), ny, nx, TL, TR, _p(A, _dp), _p(b, _dp))
    return A, b, K


def orc_assemble3d(solid, ks=10.0, kf=1.0, TL=0.0, TR=1.0):
    lib = _orc()
    solid = np.ascontiguousarray(solid, dtype=np.uint8)
    nz, ny, nx = solid.shape
    n = nx * ny * nz
    K, A, b = np.zeros(n), np.zeros((n, 7)), np.zeros(n)
    lib.orc_kmap3d.argtypes = [_u8p, C.c_size_t, C.c_double, C.c_double, _dp]
    lib.orc_kmap3d(_p(solid, _u8p), n, ks, kf, _p(K, _dp))
    lib.orc_assemble3d.argtypes = [_dp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, _dp, _dp]
    lib.orc_assemble3d(_p(K, _dp), nx, ny, nz, TL, TR, _p(A, _dp), _p(b, _dp))
    return A, b, K


def orc_apply(A, shape, x):
    """y = A x with an assembled reference-layout matrix; shape = (nz, ny, nx)."""
    lib = _orc()
    nz, ny, nx = shape
    A = np.ascontiguousarray(A)
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.zeros(nx * ny * nz)
    lib.orc_apply.argtypes = [_dp, C.c_int, C.c_int, C.c_int, C.c_int, _dp, _dp]
    lib.orc_apply(_p(A, _dp), A.shape[1], nx, ny, nz, _p(x, _dp), _p(y, _dp))
    return y


def orc_solve2d(img, amp=(1, 1), ks=10.0, kf=1.0, TL=0.0, TR=1.0, tol=1e-5, max_iter=500000):
    """Restated reference 2D path with the reference's own Gauss-Seidel + stopping rule."""
    lib = _orc()
    img = np.ascontiguousarray(img, dtype=np.uint8)
    H, W = img.shape
    nx, ny = W * amp[0], H * amp[1]
    T, qL, qR, out = np.zeros((ny, nx)), np.zeros(ny), np.zeros(ny), np.zeros(5)
    lib.orc_solve2d.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double,
                                C.c_double, C.c_double, C.c_double, C.c_long, _dp, _dp, _dp, _dp]
    lib.orc_solve2d(_p(img, _u8p), W, H, amp[0], amp[1], ks, kf, TL, TR, tol, max_iter, _p(T, _dp),
                    _p(qL, _dp), _p(qR, _dp), _p(out, _dp))
    return dict(keff=out[0], Q1=out[1], Q2=out[2], iters=int(out[3]), residual=out[4], T=T, qL=qL, qR=qR)


def orc_solve3d(solid, ks=10.0, kf=1.0, TL=0.0, TR=1.0, tol=1e-5, max_iter=500000):
    lib = _orc()
    solid = np.ascontiguousarray(solid, dtype=np.uint8)
    nz, ny, nx = solid.shape
    T, out = np.zeros((nz, ny, nx)), np.zeros(4)
    lib.orc_solve3d.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double,
                                C.c_double, C.c_double, C.c_long, _dp, _dp]
    lib.orc_solve3d(_p(solid, _u8p), nx, ny, nz, ks, kf, TL, TR, tol, max_iter, _p(T, _dp), _p(out, _dp))
    return dict(keff=out[0], Q1=out[1], Q2=out[2], iters=int(out[3]), T=T)


def orc_discrete(s, dim=None, ks=10.0, kf=1.0, TL=0.0, TR=1.0, eps=1e-12, max_iter=100000):
    """Discrete solution of the reference's own linear system (tightened-reference limit).

    dim=2: ``s`` is a u8 image [ny,nx], pixel<150 = fluid.  dim=3: u8 [nz,ny,nx], 1 = solid."""
    lib = _orc()
    s = np.ascontiguousarray(s, dtype=np.uint8)
    dim = dim or s.ndim
    shape = s.shape if dim == 3 else (1,) + s.shape
    nz, ny, nx = shape
    T, out = np.zeros(shape), np.zeros(5)
    lib.orc_discrete.argtypes = [C.c_int, _u8p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double,
                                 C.c_double, C.c_double, C.c_double, C.c_long, _dp, _dp]
    lib.orc_discrete(dim, _p(s, _u8p), nx, ny, nz, ks, kf, TL, TR, eps, max_iter, _p(T, _dp), _p(out, _dp))
    return dict(keff=out[0], Q1=out[1], Q2=out[2], iters=int(out[3]), relres=out[4],
                T=T.reshape(s.shape))


def orc_gs3d_sweeps(solid, n_sweeps, ks=10.0, kf=1.0, TL=0.0, TR=1.0):
    lib = _orc()
    solid = np.ascontiguousarray(solid, dtype=np.uint8)
    nz, ny, nx = solid.shape
    lib.orc_gs3d_sweeps.restype = C.c_long
    lib.orc_gs3d_sweeps.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double,
                                    C.c_double, C.c_long]
    return lib.orc_gs3d_sweeps(_p(solid, _u8p), nx, ny, nz, ks, kf, TL, TR, n_sweeps)
