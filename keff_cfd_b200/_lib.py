"""ctypes binding of libkeffb200.so (the C ABI declared in include/keffb200.h).

There is no CPU path: if the CUDA library is missing, or no B200 is visible when a solve is
requested, the call raises -- it never falls back to anything else.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "csrc", "libkeffb200.so")


class Params(C.Structure):
    """keffb200_params (include/keffb200.h)."""
    _fields_ = [("ks", C.c_double), ("kf", C.c_double), ("tl", C.c_double), ("tr", C.c_double),
                ("convergence", C.c_double), ("max_iter", C.c_long), ("rel_residual", C.c_double),
                ("flux_balance", C.c_double), ("check_every", C.c_int), ("flags", C.c_int)]


class Result(C.Structure):
    """keffb200_result (include/keffb200.h)."""
    _fields_ = [("keff", C.c_double), ("q_left", C.c_double), ("q_right", C.c_double),
                ("rel_residual", C.c_double), ("energy_residual", C.c_double), ("iters", C.c_long),
                ("status", C.c_int), ("gpu_ms", C.c_float)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


F_NO_TMA = 1
F_ZERO_INIT = 2
F_BATCH_GLOBAL = 4
F_NO_MG = 8
F_BATCH_ADDITIVE = 16
COMM_ID_BYTES = 128

# every symbol include/keffb200.h declares (tests check the built library exports each of them)
EXPORTS = [
    "keffb200_init", "keffb200_shutdown", "keffb200_last_error", "keffb200_version",
    "keffb200_default_params", "keffb200_launch_count", "keffb200_solve2d", "keffb200_solve2d_batch",
    "keffb200_solve3d", "keffb200_solve3d_dev", "keffb200_solve2d_batch_dev", "keffb200_comm_id",
    "keffb200_dist_init", "keffb200_dist_shutdown", "keffb200_solve3d_slab_dev", "keffb200_apply",
    "keffb200_flux", "keffb200_bench_apply_dev", "keffb200_gen_bernoulli_dev", "keffb200_precond",
    "keffb200_timer_start", "keffb200_timer_stop", "keffb200_init_multi", "keffb200_device_count",
    "keffb200_set_slab_min_cells", "keffb200_dist_peer_memory", "keffb200_wait_stream", "keffb200_flux_field",
    "keffb200_interpolate", "keffb200_host_alloc", "keffb200_host_free",
]

_lib = None


class KeffB200Error(RuntimeError):
    pass


def lib():
    """Load libkeffb200.so (built in-tree by __graft_entry__.build() / csrc/Makefile)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise KeffB200Error(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(make -C keff_cfd_b200/csrc). There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        vp, dp, u8p = C.c_void_p, C.c_void_p, C.c_void_p  # raw addresses: host or device
        PP, RP = C.POINTER(Params), C.POINTER(Result)
        L.keffb200_last_error.restype = C.c_char_p
        L.keffb200_version.restype = C.c_char_p
        L.keffb200_launch_count.restype = C.c_long
        L.keffb200_init.argtypes = [C.c_int]
        L.keffb200_init_multi.argtypes = [C.c_int]
        L.keffb200_set_slab_min_cells.argtypes = [C.c_long]
        L.keffb200_set_slab_min_cells.restype = None
        L.keffb200_wait_stream.argtypes = [C.c_void_p]
        L.keffb200_flux_field.argtypes = [u8p, C.c_int, C.c_int, C.c_int, PP, dp, dp]
        L.keffb200_interpolate.argtypes = [dp, C.c_int, C.c_int, C.c_int, dp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double]
        L.keffb200_host_alloc.argtypes = [C.c_size_t]
        L.keffb200_host_alloc.restype = C.c_void_p
        L.keffb200_host_free.argtypes = [C.c_void_p]
        L.keffb200_host_free.restype = None
        L.keffb200_default_params.argtypes = [PP]
        L.keffb200_solve2d.argtypes = [u8p, C.c_int, C.c_int, PP, dp, dp, dp, dp, RP]
        L.keffb200_solve2d_batch.argtypes = [u8p, C.c_int, C.c_int, C.c_int, PP, dp, RP]
        L.keffb200_solve3d.argtypes = [u8p, C.c_int, C.c_int, C.c_int, PP, dp, dp, dp, dp, RP]
        L.keffb200_solve3d_dev.argtypes = [u8p, C.c_int, C.c_int, C.c_int, PP, dp, dp, RP]
        L.keffb200_solve2d_batch_dev.argtypes = [u8p, C.c_int, C.c_int, C.c_int, PP, dp, RP]
        L.keffb200_comm_id.argtypes = [vp]
        L.keffb200_dist_init.argtypes = [C.c_int, C.c_int, vp]
        L.keffb200_solve3d_slab_dev.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, PP, dp, RP]
        L.keffb200_apply.argtypes = [u8p, C.c_int, C.c_int, C.c_int, PP, dp, dp, dp]
        L.keffb200_flux.argtypes = [u8p, C.c_int, C.c_int, C.c_int, PP, dp, dp, dp, RP]
        L.keffb200_precond.argtypes = [u8p, C.c_int, C.c_int, C.c_int, PP, dp, dp]
        L.keffb200_bench_apply_dev.argtypes = [u8p, C.c_int, C.c_int, C.c_int, PP, C.c_int,
                                               C.POINTER(C.c_float)]
        L.keffb200_timer_stop.argtypes = [C.POINTER(C.c_float)]
        L.keffb200_gen_bernoulli_dev.argtypes = [u8p, C.c_size_t, C.c_ulonglong, C.c_ulonglong, C.c_double]
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise KeffB200Error(f"keffb200 error {rc}: {lib().keffb200_last_error().decode()}")


_inited = None


def init(device=0):
    """Bind this process to one GPU (one process per GPU)."""
    global _inited
    if _inited != device:
        check(lib().keffb200_init(int(device)))
        _inited = device


def init_multi(n_gpus=0):
    """Bind this process to GPUs 0..n-1 (0 = all visible): host-buffer batch and volume solves fan out over them."""
    global _inited
    check(lib().keffb200_init_multi(int(n_gpus)))
    _inited = 0
    return lib().keffb200_device_count()


def order_after_torch(tensor=None):
    """Device-buffer calls read their inputs on the library's own stream: make it wait for what torch has queued
    on its current stream (the kernels still writing those tensors)."""
    try:
        import torch
    except ImportError:
        return
    if torch.cuda.is_available() and _inited is not None:
        dev = tensor.device if tensor is not None and getattr(tensor, "is_cuda", False) else torch.device("cuda", _inited)
        check(lib().keffb200_wait_stream(C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))


def ensure(device=None, tensor=None):
    """One process is bound to ONE GPU.  device=None keeps the binding made by init() (binding to GPU 0 if there
    is none yet); an explicit device re-binds only if it differs.  A CUDA tensor argument must live on the bound
    GPU -- a kernel launched on another device would fault on its address (or, with peer access enabled, silently
    run every rank's work on one GPU)."""
    if device is None:
        device = _inited if _inited is not None else (tensor.device.index if tensor is not None and tensor.is_cuda else 0)
    init(device)
    if tensor is not None and getattr(tensor, "is_cuda", False) and tensor.device.index != _inited:
        raise KeffB200Error(f"tensor lives on cuda:{tensor.device.index} but this process is bound to cuda:{_inited}")
    return _inited


def shutdown():
    global _inited
    if _lib is not None:
        _lib.keffb200_shutdown()
    _inited = None
    try:
        from . import dist
        dist._comm_ready = False
    except Exception:
        pass
