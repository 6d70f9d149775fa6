"""torchrun --nproc-per-node N scratch/dbg_slabT.py [pre512] [reps]: replays bench.py's slab_parity sequence and reports,
per slab, how far the gathered temperature field is from the single-GPU field."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import keff_cfd_b200 as K
from keff_cfd_b200 import api, dist as D
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl")
pre = int(sys.argv[1]) if len(sys.argv) > 1 else 1
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
n = 256
K.init(local)
p = api.default_params()
def barrier():
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
if pre and rank == 0:
    vol = api.bernoulli_volume_dev((512, 512, 512), 0.5, 1234, device=local)
    r3 = api.solve3d(vol, p, want_T=False, device=local)
    print("pre 512:", r3["keff"], r3["iters"], flush=True)
    del vol
torch.cuda.empty_cache()
D.init_comm(rank, world, local)
z0, nzl = D.slab_extent(n, rank, world)
loc = api.bernoulli_volume_dev((nzl, n, n), 0.5, 1234, z0=z0, device=local)
halo = D.exchange_structure_halos(loc, rank, world)
del loc
for k in range(reps):
    barrier()
    r = D.solve3d_slab(halo, z0, n, p, want_T=True)
    parts = [torch.empty_like(r["T"]) for _ in range(world)]
    dist.all_gather(parts, r["T"])
    if rank == 0:
        print(f"rep {k}: keff {r['keff']!r} iters {r['iters']} relres {r['rel_residual']:.3e}", flush=True)
    if k == 0:
        parts0 = [q.clone() for q in parts]
if rank == 0:
    full = api.bernoulli_volume_dev((n, n, n), 0.5, 1234, device=local)
    s = api.solve3d(full, p, device=local)
    print("single:", s["keff"], s["iters"])
    for name, pp in (("first", parts0), ("last", parts)):
        d = [float((pp[q] - s["T"][q * nzl:(q + 1) * nzl]).abs().max().item()) for q in range(world)]
        print(name, "per-slab T maxdiff:", ["%.2e" % v for v in d], flush=True)
        bad = [q for q in range(world) if d[q] > 1e-6]
        for q in bad[:2]:
            t = pp[q]
            print("  slab", q, "min/max", float(t.min()), float(t.max()), "nonfinite", int((~torch.isfinite(t)).sum()), "zeros", int((t == 0).sum()), "of", t.numel(),
                  "per-plane maxdiff", ["%.1e" % float((t[z] - s["T"][q * nzl + z]).abs().max()) for z in range(0, nzl, max(1, nzl // 8))])
barrier()
dist.destroy_process_group()
