#!/usr/bin/env python
"""bench.py -- the reference's headline metric on its named configuration.

metric  : converged Keff solves/s (BASELINE.json), workload = configs[1]: a batch of 10,000 synthetic
          128x128 binary 2D images per GPU (overlapping-disc "Mendeley-shaped" structures, ks:kf = 10:1,
          TL 0, TR 1).  One step = one pass of the hot path over the batch.  Weak scaling: every rank
          solves its own 10,000 images, no data-path collective (images are independent).
value   : images / s with the structures already resident in HBM (keffb200_solve2d_batch_dev).
e2e     : the same through the host-buffer C-ABI call (keffb200_solve2d_batch): pinned host structures
          copied H2D and result records copied D2H inside the timed region, every step.
extras  : "stencil3d" -- the 3D stencil kernel on a 512^3 random binary volume (configs[3]):
          Gcell-updates/s and fraction of the measured HBM roofline; "solve3d" -- one converged
          512^3 solve.  (Skipped with --no-3d.)
--impl reference : the reference's own CPU batch path (oracle/_ref: unmodified Keff2D code, OpenMP
          over images, SerialGS each, shipped Convergence 1e-5) on a bounded sample of the same images.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly ONE JSON line: NCCL's banner / debug lines (NCCL_DEBUG=VERSION prints "NCCL version ...") go to stderr
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

N_IMG, NPIX = 10000, 128
METRIC, UNIT = "converged Keff solves/sec", "solves/s"


def measured_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def ncu_traffic(kernel, units):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture
    (profiles/traffic.json: bytes per unit of work, measured at the bench's own launch size), else None."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[kernel]
        return float(t["dram_bytes_per_unit"]) * units
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (one background
    `nvidia-smi -lms 100` process, started before and killed after, as in the profiling recipe)."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.proc = index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def summary(self):
        rows = []
        if self.proc is not None:
            self.proc.terminate()
            try:
                out, _ = self.proc.communicate(timeout=5)
            except Exception:
                self.proc.kill()
                out = ""
            rows = [[c.strip() for c in line.split(",")] for line in out.strip().splitlines()]
        sm = [float(r[0]) for r in rows if len(r) == 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if len(r) == 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in rows if len(r) == 6 for n, v in zip(names, r[2:]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def make_images(count, start=0):
    from keff_cfd_b200 import structures as S
    cache = f"/tmp/keffb200_discs_{NPIX}_{start}_{count}.npy"
    if os.path.exists(cache):
        return np.load(cache)
    imgs = S.disc_images(count, NPIX, start=start)
    try:
        np.save(cache, imgs)
    except Exception:
        pass
    return imgs


def run_reference(args):
    """Reference CPU arm: rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import refbind
    if not refbind.have_ref():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref was not built (needs /root/reference)"}))
        return
    cores = os.cpu_count() or 1
    per_step = max(cores, 8)
    imgs = make_images(per_step * (args.steps + args.warmup))
    gray = np.where(imgs == 1, 255, 0).astype(np.uint8)
    t_steps = []
    for s in range(args.warmup + args.steps):
        chunk = gray[s * per_step:(s + 1) * per_step]
        t0 = time.perf_counter()
        refbind.ref2d_batch(chunk, tol=1e-5, threads=cores)
        if s >= args.warmup:
            t_steps.append(time.perf_counter() - t0)
    total = sum(t_steps)
    v = per_step * len(t_steps) / total
    sample = f"{per_step} images/step x {len(t_steps)} steps of the same disc images, Convergence 1e-5 (shipped), OpenMP over images"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(t_steps), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"bounded sample of configs[1]: {per_step} of the 10000 synthetic 128x128 images per step"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--images", type=int, default=N_IMG)
    ap.add_argument("--no-3d", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    # stdout must carry exactly ONE JSON line: whatever the libraries print while the job runs (NCCL's version
    # banner, for one) goes to stderr; the real stdout is restored for the result line
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    import keff_cfd_b200 as K
    from keff_cfd_b200 import _lib, api

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    K.init(local)
    lib = K.lib()
    dev = torch.device("cuda", local)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    n_img = args.images
    imgs = make_images(n_img, start=rank * n_img)                  # every rank its own 10,000 images
    host = torch.from_numpy(imgs).pin_memory()                     # pinned host structures (e2e leg)
    devm = host.to(dev)                                            # resident structures (value leg)
    p = api.default_params()
    cells = NPIX * NPIX

    # ---- value: structures resident in HBM ----------------------------------------------------
    for _ in range(args.warmup):
        res = api.solve2d_batch(devm, p)
    import ctypes

    def timed(fn):
        """K steps bracketed by barrier + synchronize; device time from CUDA events recorded on the
        library's compute stream (keffb200_timer_*), wall time beside it; both max over ranks."""
        barrier()
        ms = ctypes.c_float()
        t0 = time.perf_counter()
        _lib.check(lib.keffb200_timer_start())
        for _ in range(args.steps):
            fn()
        _lib.check(lib.keffb200_timer_stop(ctypes.byref(ms)))
        barrier()
        wall = time.perf_counter() - t0
        return max_over_ranks(ms.value * 1e-3), max_over_ranks(wall)

    l0 = lib.keffb200_launch_count()
    clk = ClockSampler(local)
    clk.start()
    kernel_ms = 0.0
    out_res = []

    def step_value():
        nonlocal kernel_ms
        r = api.solve2d_batch(devm, p)
        kernel_ms += float(r["gpu_ms"][0])                         # the launch alone (CUDA events inside the call)
        out_res[:] = [r]

    dt, wall_v = timed(step_value)
    res = out_res[0]
    clocks = clk.summary()
    launches = lib.keffb200_launch_count() - l0
    value = world * n_img * args.steps / dt
    iters = np.asarray(res["iters"])
    ok = bool((res["status"] == 0).all())
    # cell-updates (one evaluation of a cell's stencil row, operator or relaxation).  Per CG iteration the
    # on-chip multigrid kernel makes three passes over the image (A p, the restriction's residual, the
    # post-smoothing sweep) and four over every coarse level (two sweeps per side; 1/4 + 1/16 + 1/64 of the cells);
    # the residual evaluations of the reliable updates are not counted (the kernel does not report how many it made).
    PASSES = 3.0 + 4.0 * (1 / 4 + 1 / 16 + 1 / 64)
    updates_per_step = float((iters * PASSES + 2).sum()) * cells
    k_ms = kernel_ms / args.steps
    peak, peak_kind = measured_peak()
    achieved = 16.0 * updates_per_step / (k_ms * 1e-3) / 1e9

    # ---- e2e: host buffers through the reference-facing C-ABI call ----------------------------------
    harr = host.numpy()
    for _ in range(min(args.warmup, 2)):
        api.solve2d_batch(harr, p)
    out_h = []
    dt_e, wall_e = timed(lambda: out_h.__setitem__(slice(None), [api.solve2d_batch(harr, p)]))
    res_h = out_h[0]
    e2e = world * n_img * args.steps / dt_e
    keff_mean = float(res_h["keff"].mean())

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "wall_ms_per_step": 1e3 * wall_v / args.steps,
        "timing": "CUDA events on the library's compute stream around the K steps, max over ranks",
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"configs[1]: {n_img} synthetic {NPIX}x{NPIX} binary images per GPU (overlapping discs, "
                               "ks:kf 10:1), TRUE relative residual 1e-8 + flux balance 1e-6, on-chip multigrid-preconditioned mixed-"
                               "precision CG with FP64 reliable updates",
                   "images_per_gpu": n_img, "l2": "inputs larger than L2 (164 MB of structures per step; every CTA "
                   "rewrites its CG vectors for each image)", "all_converged": bool(ok),
                   "mean_iters": float(iters.mean()), "mean_keff": keff_mean},
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(n_img * cells),
                "d2h_bytes_per_step": int(n_img * 48),
                "ms_per_step": 1e3 * dt_e / args.steps, "wall_ms_per_step": 1e3 * wall_e / args.steps},
        "gpu_launches": int(launches), "clocks": clocks,
        "roofline": {"kernel": "k_batch2d_mg", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": ncu_traffic("k_batch2d_mg", n_img), "peak_kind": peak_kind,
                     "algorithmic_bytes_per_launch": 16.0 * updates_per_step,
                     "note": "HBM-equivalent: 16 B x cell-updates / kernel time (4.31 stencil passes per CG iteration, see "
                             "DESIGN 3).  The whole solve of an image runs inside one SM (shared memory, registers, an "
                             "L2-resident slice), so this kernel is issue/latency bound, not HBM bound -- its real DRAM "
                             "traffic is `traffic` -- and the figure is a work rate, not a bandwidth claim (SURVEY 8d, "
                             "config-2 caveat); the HBM-bound kernel of the path is reported under stencil3d",
                     "gcell_updates_per_s": updates_per_step / (k_ms * 1e-3) / 1e9},
    }

    if rank == 0 and not args.no_3d:
        n3 = 512
        vol = api.bernoulli_volume_dev((n3, n3, n3), 0.5, 1234, device=local)
        ms = api.bench_apply(vol, p, reps=20, device=local)
        gb = 16.0 * n3 ** 3 / (ms * 1e-3) / 1e9
        out["stencil3d"] = {"workload": "configs[3]: 512^3 random binary volume, ks:kf 10:1", "kernel": "k_stencil_tma",
                            "ms_per_sweep": ms, "gcell_updates_per_s": n3 ** 3 / (ms * 1e-3) / 1e9, "bound": "hbm",
                            "achieved": gb, "peak": peak, "unit": "GB/s", "frac": gb / peak,
                            "algorithmic_bytes_per_launch": 16.0 * n3 ** 3,
                            "traffic": ncu_traffic("k_stencil_tma_apply", n3 ** 3)}
        api.solve3d(vol, p, want_T=False, device=local)               # first call allocates the workspace
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r3 = api.solve3d(vol, p, want_T=False, device=local)
        torch.cuda.synchronize()
        out["solve3d"] = {"workload": "configs[3]", "solver": "CG preconditioned by a geometric multigrid V-cycle "
                          "(FP32 cycle, FP64 CG vectors)", "seconds": time.perf_counter() - t0,
                          "gpu_ms": r3["gpu_ms"], "solves_per_s": 1e3 / r3["gpu_ms"], "iters": r3["iters"],
                          "ms_per_iter": r3["gpu_ms"] / max(r3["iters"], 1), "keff": r3["keff"],
                          "rel_residual": r3["rel_residual"], "status": r3["status"]}
        del vol

    if world > 1 and not args.no_3d:
        # configs[4]-shaped weak scaling of the z-slab solve: 1024 x 1024 x 128 planes per GPU (134 M cells per
        # GPU, the same as the 512^3 single-GPU case; at 8 GPUs this IS configs[4], the 1024^3 volume)
        from keff_cfd_b200 import dist as KD
        del devm
        torch.cuda.empty_cache()
        KD.init_comm(rank, world, local)
        nxs, nzg = 1024, 128 * world
        z0, nzl = KD.slab_extent(nzg, rank, world)
        loc = api.bernoulli_volume_dev((nzl, nxs, nxs), 0.5, 1234, z0=z0, device=local)
        halo = KD.exchange_structure_halos(loc, rank, world)
        del loc
        ps = api.default_params()
        for _ in range(2):                                          # first call allocates the workspace
            barrier()
            r3 = KD.solve3d_slab(halo, z0, nzg, ps)
        ms3 = max_over_ranks(float(r3["gpu_ms"]))
        if rank == 0:
            out["slab3d"] = {"workload": f"1024x1024x{nzg} random binary volume in {world} z-slabs of 128 planes "
                             "(configs[4] at 8 GPUs), ks:kf 10:1, halo planes + scalar all-reduces over NCCL",
                             "gpu_ms": ms3, "solves_per_s": 1e3 / ms3, "iters": r3["iters"],
                             "ms_per_iter": ms3 / max(r3["iters"], 1), "keff": r3["keff"],
                             "rel_residual": r3["rel_residual"], "status": r3["status"],
                             "gcell_iters_per_s": nxs * nxs * nzg * r3["iters"] / (ms3 * 1e-3) / 1e9,
                             "scaling": "weak (compare gpu_ms with solve3d.gpu_ms of the 1-GPU line: 512^3, same cells per GPU)"}
        del halo

    if rank == 0 and not args.no_cpu:
        from oracle import refbind
        cores = os.cpu_count() or 1
        k = max(cores, 8)
        gray = np.where(imgs[:k] == 1, 255, 0).astype(np.uint8)
        t0 = time.perf_counter()
        if refbind.have_ref():
            ref = refbind.ref2d_batch(gray, tol=1e-5, threads=cores)
            kind = "reference"
        else:
            ref = np.array([[refbind.orc_solve2d(g, tol=1e-5)[q] for q in ("keff", "Q1", "Q2", "iters")] for g in gray])
            kind, cores = "port", 1
        tc = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": k / tc, "unit": UNIT, "cores": cores, "kind": kind,
                               "sample": f"first {k} images of the batch, reference batch mode (OpenMP over images, "
                                         f"SerialGS each) at the shipped Convergence 1e-5; mean sweeps {ref[:, 3].mean():.0f}",
                               "max_keff_gap_vs_gpu": float(np.max(np.abs(ref[:, 0] - res_h["keff"][:k]) / res_h["keff"][:k])),
                               "gap_note": "the reference stops on its own loose test (Keff changed < 1e-5 between two checks), "
                                           "1e-3..1e-2 away from the discrete solution on these images (SURVEY 8c); the GPU "
                                           "results are at TRUE relative residual 1e-8.  Parity proper is tests/test_gpu_parity.py "
                                           "(tight reference, Keff within 1e-5)"}
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    if rank == 0:
        print(json.dumps(out), flush=True)
    os.dup2(2, 1)  # anything printed during teardown goes to stderr again
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
