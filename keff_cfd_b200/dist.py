"""One-process-per-GPU plumbing (torch.distributed) around the C ABI.

Partitioning follows the data (BASELINE north_star):
  * batches of independent 2D images are split across ranks with NO communication
    (the reference treats images independently too: Keff2D/keff2D.cpp:341);
  * a large 3D volume is cut into z-slabs -- z is the slowest index (Keff3D/keff3D.h:691) and the
    global z faces are adiabatic, so a slab boundary cuts only ordinary faces and each halo is one
    contiguous nx*ny plane.  The per-iteration halo exchange and the two scalar all-reduces run
    inside libkeffb200 on its own NCCL communicator; torch.distributed only carries the setup
    (communicator id, structure halo planes) and gathers results.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import Result, check, lib


def slab_extent(nz, rank, world):
    """Planes [z0, z0+nzl) owned by `rank`: contiguous, sizes differ by at most one."""
    base, rem = divmod(nz, world)
    nzl = base + (1 if rank < rem else 0)
    z0 = rank * base + min(rank, rem)
    return z0, nzl


def image_shard(n_img, rank, world):
    """Images [i0, i0+cnt) owned by `rank` (contiguous blocks, sizes differ by at most one)."""
    return slab_extent(n_img, rank, world)


def exchange_structure_halos(local, rank, world, group=None):
    """local: u8 tensor [nzl, ny, nx] (CPU for gloo, CUDA for nccl).  Returns [nzl+2, ny, nx] with the
    neighbours' boundary planes in plane 0 and plane nzl+1 (zeros at the global ends)."""
    import torch
    import torch.distributed as dist
    nzl, ny, nx = local.shape
    out = torch.zeros((nzl + 2, ny, nx), dtype=local.dtype, device=local.device)
    out[1:-1] = local
    if world == 1:
        return out
    ops = []
    lo_send, hi_send = local[0].contiguous(), local[-1].contiguous()
    lo_recv, hi_recv = torch.empty_like(lo_send), torch.empty_like(hi_send)
    if rank > 0:
        ops += [dist.P2POp(dist.isend, lo_send, rank - 1, group), dist.P2POp(dist.irecv, lo_recv, rank - 1, group)]
    if rank < world - 1:
        ops += [dist.P2POp(dist.isend, hi_send, rank + 1, group), dist.P2POp(dist.irecv, hi_recv, rank + 1, group)]
    for w in dist.batch_isend_irecv(ops):
        w.wait()
    if rank > 0:
        out[0] = lo_recv
    if rank < world - 1:
        out[-1] = hi_recv
    return out


_comm_ready = False


def init_comm(rank, world, device):
    """Create libkeffb200's NCCL communicator: rank 0 makes the id, torch.distributed broadcasts it."""
    global _comm_ready
    import torch.distributed as dist
    _lib.init(device)
    if _comm_ready:
        return
    buf = (C.c_char * _lib.COMM_ID_BYTES)()
    if world > 1:
        payload = [None]
        if rank == 0:
            check(lib().keffb200_comm_id(buf))
            payload = [bytes(buf.raw)]
        dist.broadcast_object_list(payload, src=0)
        C.memmove(buf, payload[0], _lib.COMM_ID_BYTES)
    check(lib().keffb200_dist_init(rank, world, buf))
    _comm_ready = True


def peer_memory():
    """True when the slab group exchanges halos and reductions through peer memory (NVLink stores), False on NCCL."""
    return bool(lib().keffb200_dist_peer_memory())


def solve3d_slab(solid_halo_dev, z0, nz_global, params, want_T=False):
    """Collective z-slab solve.  solid_halo_dev: CUDA u8 [nzl+2, ny, nx] (from
    exchange_structure_halos).  Every rank gets the same result record."""
    import torch
    nzl, ny, nx = solid_halo_dev.shape[0] - 2, solid_halo_dev.shape[1], solid_halo_dev.shape[2]
    T = torch.empty((nzl, ny, nx), dtype=torch.float64, device=solid_halo_dev.device) if want_T else None
    r = Result()
    _lib.order_after_torch(solid_halo_dev)   # exchange_structure_halos fills the halo planes with asynchronous torch ops
    check(lib().keffb200_solve3d_slab_dev(solid_halo_dev.data_ptr(), nx, ny, nzl, z0, nz_global, C.byref(params),
                                          None if T is None else T.data_ptr(), C.byref(r)))
    d = r.as_dict()
    d["T"] = T
    return d


def gather_batch_results(local_results, rank, world):
    """Per-image records of every rank, in global image order, on rank 0 (None elsewhere)."""
    import torch.distributed as dist
    if world == 1:
        return local_results
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(local_results, gathered, dst=0)
    if rank != 0:
        return None
    return [r for part in gathered for r in part]


def keff_table(results):
    return np.array([[r["keff"], r["q_left"], r["q_right"], r["iters"]] for r in results])
