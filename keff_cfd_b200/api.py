"""Host-side mirror of the reference's interface for the hot path, over the C ABI.

The reference has no library API: each program reads ``input.txt`` into an ``options`` struct
(Keff2D/keff2D.h:15-36, Keff3D/keff3D.h:13-38), builds the structure, and calls the solver.  This
module keeps those names and meanings (``Options`` fields are the struct's fields, ``read_input_file``
is ``readInputFile``) and forwards the numerical work to libkeffb200.so.  Nothing here computes.
"""
import ctypes as C
import dataclasses

import numpy as np

from . import _lib
from ._lib import Params, Result, check, lib


@dataclasses.dataclass
class Options:
    """The reference ``options`` struct; defaults are the shipped input.txt values."""
    TCsolid: float = 10.0
    TCfluid: float = 1.0
    Width: int = 0
    Height: int = 0
    Depth: int = 0
    MeshIncreaseX: int = 1
    MeshIncreaseY: int = 1
    MeshIncreaseZ: int = 1
    TempLeft: float = 0.0
    TempRight: float = 1.0
    MAX_ITER: int = 500000
    ConvergeCriteria: float = 1e-5
    inputFilename: str = ""
    outputFilename: str = "output.csv"
    printTmap: int = 0
    TMapName: str = "TMAP.csv"
    printQmap: int = 0
    QMapName: str = "QMAP.csv"
    verbose: int = 0
    numCores: int = 1
    MeshFlag: int = 0
    MaxMesh: int = 1
    BatchFlag: int = 0
    NumImg: int = 0


_KEYS = {  # input.txt key -> (field, type); Keff2D/keff2D.h:142-205, Keff3D/keff3D.h:145-220
    "ks:": ("TCsolid", float), "kf:": ("TCfluid", float), "Width:": ("Width", int), "Height:": ("Height", int),
    "Depth:": ("Depth", int), "MeshAmpX:": ("MeshIncreaseX", int), "MeshAmpY:": ("MeshIncreaseY", int),
    "MeshAmpZ:": ("MeshIncreaseZ", int), "InputName:": ("inputFilename", str), "TR:": ("TempRight", float),
    "TL:": ("TempLeft", float), "OutputName:": ("outputFilename", str), "printTMap:": ("printTmap", int),
    "TMapName:": ("TMapName", str), "printQMap:": ("printQmap", int), "QMapName:": ("QMapName", str),
    "Convergence:": ("ConvergeCriteria", float), "MaxIter:": ("MAX_ITER", int), "Verbose:": ("verbose", int),
    "NumCores:": ("numCores", int), "MeshFlag:": ("MeshFlag", int), "MaxMesh:": ("MaxMesh", int),
    "RunBatch:": ("BatchFlag", int), "NumImages:": ("NumImg", int),
}


def read_input_file(path):
    """``readInputFile``: case-sensitive ``key: value`` lines, order-free, unknown lines ignored."""
    o = Options()
    with open(path) as f:
        for line in f:
            tok = line.split()
            if len(tok) < 2 or tok[0] not in _KEYS:
                continue
            field, typ = _KEYS[tok[0]]
            setattr(o, field, tok[1] if typ is str else typ(float(tok[1])))
    if o.numCores < 1:
        o.numCores = 1
    return o


def default_params(**kw):
    p = Params()
    lib().keffb200_default_params(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise TypeError(f"unknown solver parameter {k!r}")
        setattr(p, k, v)
    return p


def params_from_options(o, **kw):
    return default_params(ks=o.TCsolid, kf=o.TCfluid, tl=o.TempLeft, tr=o.TempRight,
                          convergence=o.ConvergeCriteria, max_iter=o.MAX_ITER, **kw)


def structure_from_image(img, amp_x=1, amp_y=1):
    """2D structure rule of the reference: pixel < 150 is fluid, else solid; MeshAmp is an
    integer-division replicate (Keff2D/keff2D.cpp:108-121).  Returns u8 [ny, nx], 1 = solid."""
    img = np.asarray(img)
    solid = (img >= 150).astype(np.uint8)
    if amp_x != 1 or amp_y != 1:
        solid = np.repeat(np.repeat(solid, amp_y, axis=0), amp_x, axis=1)
    return np.ascontiguousarray(solid)


def structure_from_csv(path, width, height, depth):
    """3D structure rule: every ``x,y,z`` line marks a solid voxel (Keff3D/keff3D.h:269-277)."""
    xyz = np.loadtxt(path, delimiter=",", dtype=np.int64, ndmin=2)
    s = np.zeros((depth, height, width), dtype=np.uint8)
    s[xyz[:, 2], xyz[:, 1], xyz[:, 0]] = 1
    return s


def _addr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return a.data_ptr()  # torch tensor (host pinned or device)


def _is_cuda(a):
    return hasattr(a, "is_cuda") and a.is_cuda


def _res(r, **extra):
    d = r.as_dict()
    d.update(extra)
    return d


def solve3d(solid, params=None, T_init=None, want_T=True, want_q=False, device=None):
    """One volume.  ``solid``: u8 [nz, ny, nx] numpy array / pinned torch tensor (host path: the copy
    to the GPU is part of the call) or CUDA torch tensor (device path)."""
    device = _lib.ensure(device, solid if _is_cuda(solid) else None)
    p = params or default_params()
    nz, ny, nx = solid.shape
    r = Result()
    if _is_cuda(solid):
        import torch
        T = torch.empty((nz, ny, nx), dtype=torch.float64, device=solid.device) if want_T else None
        _lib.order_after_torch(solid)
        check(lib().keffb200_solve3d_dev(_addr(solid), nx, ny, nz, C.byref(p), _addr(T_init), _addr(T), C.byref(r)))
        return _res(r, T=T)
    solid = np.ascontiguousarray(solid, dtype=np.uint8) if isinstance(solid, np.ndarray) else solid
    T = np.empty((nz, ny, nx)) if want_T else None
    qL = np.empty((nz, ny)) if want_q else None
    qR = np.empty((nz, ny)) if want_q else None
    if T_init is not None:
        T_init = np.ascontiguousarray(T_init, dtype=np.float64)
    check(lib().keffb200_solve3d(_addr(solid), nx, ny, nz, C.byref(p), _addr(T_init), _addr(T), _addr(qL), _addr(qR),
                                 C.byref(r)))
    return _res(r, T=T, qL=qL, qR=qR)


def solve2d(solid, params=None, T_init=None, want_T=True, want_q=False, device=None):
    """One image: u8 [ny, nx], 1 = solid."""
    out = solve3d(solid.reshape((1,) + tuple(solid.shape)), params, None if T_init is None else T_init.reshape(
        (1,) + tuple(T_init.shape)), want_T, want_q, device)
    for k in ("T", "qL", "qR"):
        if out.get(k) is not None:
            out[k] = out[k].reshape(out[k].shape[1:]) if k == "T" else out[k].reshape(-1)
    return out


def solve2d_batch(solid, params=None, want_T=False, device=None):
    """A batch of images u8 [n_img, ny, nx]; returns the per-image result records as a numpy
    structured array (res[i]["keff"], res["iters"], ...), or a list of dicts that also carry the
    temperature maps under "T" when want_T."""
    device = _lib.ensure(device, solid if _is_cuda(solid) else None)
    p = params or default_params()
    n, ny, nx = solid.shape
    res = (Result * n)()
    if _is_cuda(solid):
        import torch
        T = torch.empty((n, ny, nx), dtype=torch.float64, device=solid.device) if want_T else None
        _lib.order_after_torch(solid)
        check(lib().keffb200_solve2d_batch_dev(_addr(solid), n, nx, ny, C.byref(p), _addr(T), res))
    else:
        solid = np.ascontiguousarray(solid, dtype=np.uint8) if isinstance(solid, np.ndarray) else solid
        T = np.empty((n, ny, nx)) if want_T else None
        check(lib().keffb200_solve2d_batch(_addr(solid), n, nx, ny, C.byref(p), _addr(T), res))
    # zero-copy view of the C result records: a structured array indexable as res[i]["keff"] / res["keff"]
    table = np.ctypeslib.as_array(res).copy()
    if not want_T:
        return table
    out = [{k: table[k][i].item() for k in table.dtype.names} for i in range(n)]
    for i, o in enumerate(out):
        o["T"] = T[i]
    return out


def apply_operator(solid, x, params=None, want_b=False, device=None):
    """y = A x (and b) for the matrix-free form of the reference's assembled system."""
    device = _lib.ensure(device)
    p = params or default_params()
    solid = np.ascontiguousarray(solid, dtype=np.uint8)
    if solid.ndim == 2:
        solid = solid[None]
    nz, ny, nx = solid.shape
    x = np.ascontiguousarray(x, dtype=np.float64).reshape(nz, ny, nx)
    y = np.empty_like(x)
    b = np.empty_like(x) if want_b else None
    check(lib().keffb200_apply(_addr(solid), nx, ny, nz, C.byref(p), _addr(x), _addr(y), _addr(b)))
    return (y, b) if want_b else y


def precondition(solid, r, params=None, device=None):
    """z = M r: one multigrid V-cycle of the grid solver's preconditioner (test hook)."""
    device = _lib.ensure(device)
    p = params or default_params()
    solid = np.ascontiguousarray(solid, dtype=np.uint8)
    if solid.ndim == 2:
        solid = solid[None]
    nz, ny, nx = solid.shape
    r = np.ascontiguousarray(r, dtype=np.float64).reshape(nz, ny, nx)
    z = np.empty_like(r)
    check(lib().keffb200_precond(_addr(solid), nx, ny, nz, C.byref(p), _addr(r), _addr(z)))
    return z


def boundary_flux(solid, T, params=None, device=None):
    """Flux integration + energy residual of a given field (Keff2D/keff2D.cpp:151-170, keff2D.h:384)."""
    device = _lib.ensure(device)
    p = params or default_params()
    solid = np.ascontiguousarray(solid, dtype=np.uint8)
    if solid.ndim == 2:
        solid = solid[None]
    nz, ny, nx = solid.shape
    T = np.ascontiguousarray(T, dtype=np.float64).reshape(nz, ny, nx)
    qL, qR = np.empty((nz, ny)), np.empty((nz, ny))
    r = Result()
    check(lib().keffb200_flux(_addr(solid), nx, ny, nz, C.byref(p), _addr(T), _addr(qL), _addr(qR), C.byref(r)))
    return _res(r, qL=qL, qR=qR)


def flux_field(solid, T, params=None, device=None):
    """x-direction heat-flux field behind the Q-maps: Qx[z, y, 0..nx] (printQMAP, Keff2D/keff2D.h:427-479)."""
    device = _lib.ensure(device)
    p = params or default_params()
    solid = np.ascontiguousarray(solid, dtype=np.uint8)
    squeeze = solid.ndim == 2
    if squeeze:
        solid = solid[None]
    nz, ny, nx = solid.shape
    T = np.ascontiguousarray(T, dtype=np.float64).reshape(nz, ny, nx)
    Q = np.empty((nz, ny, nx + 1))
    check(lib().keffb200_flux_field(_addr(solid), nx, ny, nz, C.byref(p), _addr(T), _addr(Q)))
    return Q[0] if squeeze else Q


def interpolate(T, shape, tl=0.0, tr=1.0, device=None):
    """Warm start for a finer mesh: (tri)linear interpolation of T onto `shape`, walls as nodes along x."""
    device = _lib.ensure(device)
    T = np.ascontiguousarray(T, dtype=np.float64)
    src = T if T.ndim == 3 else T[None]
    dst_shape = tuple(shape) if len(shape) == 3 else (1,) + tuple(shape)
    out = np.empty(dst_shape)
    check(lib().keffb200_interpolate(_addr(src), src.shape[2], src.shape[1], src.shape[0], _addr(out), dst_shape[2], dst_shape[1],
                                     dst_shape[0], tl, tr))
    return out.reshape(shape)


def bench_apply(solid_dev, params=None, reps=20, device=None):
    """Average ms per stencil application on a device-resident structure (CUDA events)."""
    device = _lib.ensure(device, solid_dev)
    p = params or default_params()
    nz, ny, nx = solid_dev.shape
    ms = C.c_float()
    _lib.order_after_torch(solid_dev)
    check(lib().keffb200_bench_apply_dev(_addr(solid_dev), nx, ny, nz, C.byref(p), reps, C.byref(ms)))
    return ms.value


def bernoulli_volume_dev(shape, p=0.5, seed=1234, z0=0, device=None):
    """Device-side twin of structures.bernoulli_volume: CUDA u8 tensor [nz, ny, nx]."""
    import torch
    device = _lib.ensure(device)
    nz, ny, nx = shape
    out = torch.empty(shape, dtype=torch.uint8, device=f"cuda:{device}")
    # the caching allocator may hand back a block whose previous owner still has work queued on torch's stream:
    # the library's stream must run behind it before it writes the block
    _lib.order_after_torch(out)
    check(lib().keffb200_gen_bernoulli_dev(out.data_ptr(), nz * ny * nx, z0 * ny * nx, seed, p))
    return out
