"""CPU, world_size 2 over gloo: the host-side partition / halo / gather logic of the N > 1 paths."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from keff_cfd_b200 import dist as D
    from keff_cfd_b200 import structures as S
    nz, ny, nx = 11, 6, 8
    full = S.bernoulli_volume((nz, ny, nx), 0.5, 77)
    z0, nzl = D.slab_extent(nz, rank, world)
    # each rank generates only its own planes (global index offset), then trades boundary planes
    local = torch.from_numpy(S.bernoulli_volume((nzl, ny, nx), 0.5, 77, z0=z0))
    halo = D.exchange_structure_halos(local, rank, world).numpy()
    ok = np.array_equal(halo[1:-1], full[z0:z0 + nzl])
    ok &= np.array_equal(halo[0], full[z0 - 1]) if z0 > 0 else not halo[0].any()
    ok &= np.array_equal(halo[-1], full[z0 + nzl]) if z0 + nzl < nz else not halo[-1].any()
    # image sharding: contiguous shards, results gathered in global order on rank 0
    i0, cnt = D.image_shard(7, rank, world)
    mine = [{"keff": float(i), "q_left": 0.0, "q_right": 0.0, "iters": i} for i in range(i0, i0 + cnt)]
    allr = D.gather_batch_results(mine, rank, world)
    if rank == 0:
        ok &= [r["iters"] for r in allr] == list(range(7))
        ok &= D.keff_table(allr).shape == (7, 4)
    else:
        ok &= allr is None
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


def test_slab_halos_and_image_shards_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == {0: True, 1: True}
