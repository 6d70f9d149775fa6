"""Weak scaling of the z-slab 3D solve (converged, multigrid-preconditioned CG): every rank holds 134 M cells.
torchrun --nproc-per-node N scratch/scale_slab.py [iters]   (N = 1: 512^3; N > 1: 1024 x 1024 x 128 N)"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import keff_cfd_b200 as K
from keff_cfd_b200 import api, dist as D
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 500000   # cap; default: run to convergence
flags = int(sys.argv[2]) if len(sys.argv) > 2 else 0             # 8 = KEFFB200_F_NO_MG
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
D.init_comm(rank, world, local) if world > 1 else K.init(local)
nx = ny = 512 if world == 1 else 1024
nzg = 512 if world == 1 else 128 * world
z0, nzl = D.slab_extent(nzg, rank, world)
loc = api.bernoulli_volume_dev((nzl, ny, nx), 0.5, 1234, z0=z0, device=local)
halo = D.exchange_structure_halos(loc, rank, world) if world > 1 else torch.cat([torch.zeros_like(loc[:1]), loc, torch.zeros_like(loc[:1])])
del loc
p = api.default_params(max_iter=iters, check_every=64, flags=flags)
out = []
for rep in range(2):
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    t0 = time.perf_counter()
    if world > 1:
        r = D.solve3d_slab(halo, z0, nzg, p)
    else:
        import ctypes as C
        from keff_cfd_b200._lib import Result, check, lib
        r_ = Result()
        check(lib().keffb200_solve3d_slab_dev(halo.data_ptr(), nx, ny, nzl, 0, nzg, C.byref(p), None, C.byref(r_)))
        r = r_.as_dict()
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    dt = time.perf_counter() - t0
    out.append((dt, r["gpu_ms"], r["iters"]))
if rank == 0:
    dt, ms, it = out[-1]
    cells = nx * ny * nzg
    print(json.dumps({"gpus": world, "grid": [nx, ny, nzg], "iters": it, "gpu_ms": ms, "ms_per_iter": ms / it,
                      "gcell_iter_per_s": cells * it / (ms * 1e-3) / 1e9, "keff": r["keff"], "status": r["status"], "rel_residual": r["rel_residual"],
                      "solves_per_s": 1e3 / ms, "wall_s": dt, "solver": "jacobi-pcg" if flags & 8 else "mg-pcg"}))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
