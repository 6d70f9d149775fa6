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


def KD_image_shard(n_img, rank, world):
    from keff_cfd_b200 import dist as KD
    return KD.image_shard(n_img, rank, world)


def cpu_other_configs(refbind, api, p):
    """The reference CPU code on the configs it can run besides the batch, bounded:
    configs[0] = the shipped 00009.jpg at the shipped keys on one thread (deterministic), with the GPU solve of the
    same image beside it; configs[2] = a fixed number of ParallelGS3D sweeps on the shipped-size Schwarz volume."""
    out = {}
    gold = os.path.join(ROOT, "tests", "golden")
    try:
        img = np.load(os.path.join(gold, "img_00009.npy"))
        if refbind.have_ref():
            t0 = time.perf_counter()
            r = refbind.ref2d_solve(img, tol=1e-5, threads=1, want_T=False)
            tc = time.perf_counter() - t0
            kind = "reference"
        else:
            t0 = time.perf_counter()
            r = refbind.orc_solve2d(img, tol=1e-5)
            tc = time.perf_counter() - t0
            kind = "port"
        solid = api.structure_from_image(img)
        api.solve2d(solid, p, want_T=False)
        g = api.solve2d(solid, p, want_T=False)
        out["config0_00009"] = {"kind": kind, "cores": 1, "cpu_seconds": tc, "cpu_sweeps": int(r["iters"]), "cpu_keff": float(r["keff"]),
                                "cpu_mcell_updates_per_s": r["iters"] * img.size / tc / 1e6, "gpu_ms": g["gpu_ms"],
                                "gpu_iters": g["iters"], "gpu_keff": g["keff"],
                                "sample": "Keff2D 00009.jpg, shipped input.txt keys (Convergence 1e-5), one thread"}
    except Exception as e:  # a missing fixture must not cost the bench line
        out["config0_00009"] = {"error": repr(e)}
    try:
        bits = np.load(os.path.join(gold, "schwarz163_bits.npy"))
        vol = np.unpackbits(bits).reshape(128, 128, 128)
        sweeps = 200
        if refbind.have_ref():
            t0 = time.perf_counter()
            refbind.ref3d_solve(vol, tol=1e-30, max_iter=sweeps, threads=1, want_T=False)
            tc = time.perf_counter() - t0
            kind = "reference"
        else:
            t0 = time.perf_counter()
            refbind.orc_gs3d_sweeps(vol, sweeps)
            tc = time.perf_counter() - t0
            kind = "port"
        g = api.solve3d(vol, p, want_T=False)
        g = api.solve3d(vol, p, want_T=False)
        out["config2_schwarz163"] = {"kind": kind, "cores": 1, "cpu_seconds": tc, "cpu_sweeps": sweeps,
                                     "cpu_mcell_updates_per_s": sweeps * vol.size / tc / 1e6,
                                     "cpu_seconds_to_shipped_tol_extrapolated": tc / sweeps * 14600,
                                     "gpu_ms": g["gpu_ms"], "gpu_iters": g["iters"], "gpu_keff": g["keff"],
                                     "sample": f"Schwarz163 128^3: {sweeps} ParallelGS3D sweeps on one thread (the shipped tolerance "
                                               "takes 14 600 sweeps: tests/golden/ref3d_schwarz163.json)"}
    except Exception as e:
        out["config2_schwarz163"] = {"error": repr(e)}
    out["ref_gpu_program_00000"] = ref_gpu_program(api, p)
    return out


def ref_gpu_program(api, p):
    """The reference's OWN GPU program (Keff2DGPU/keff2D.cu: FP32 Jacobi through updateX_V1, host-side Residual every
    10^4 / 10^3 iterations, keff2D.cuh:726-849), unmodified, compiled for sm_100 into oracle/_ref by oracle/Makefile, on
    its shipped input (00000.jpg, MeshAmp 2 -> 256 x 256, Convergence 1e-5) -- the program this path replaces -- with
    keffb200_solve2d on the same structure beside it."""
    import shutil
    import tempfile
    exe = os.path.join(ROOT, "oracle", "_ref", "keff2DGPU")
    jpg = os.path.join(ROOT, "tests", "golden", "inputs", "00000.jpg")
    if not (os.path.exists(exe) and os.path.exists(jpg)):
        return {"unavailable": "oracle/_ref/keff2DGPU was not built (needs /root/reference and nvcc)"}
    d = tempfile.mkdtemp()
    try:
        shutil.copy(jpg, d)
        open(os.path.join(d, "input.txt"), "w").write(
            "Input File:\nks: 10\nkf: 1\nMeshAmpX: 2\nMeshAmpY: 2\nInputName: 00000.jpg\nTR: 1\nTL: 0\nOutputName: batchTest.csv\n"
            "printTMap: 0\nTMapName: T.csv\nprintQMap: 0\nQMapName: Q.csv\nConvergence: 0.00001\nMaxIter: 500000\nVerbose: 0\n"
            "NumCores: 1\nMeshFlag: 0\nMaxMesh: 1\nRunBatch: 0\nNumImages: 1\n")
        t0 = time.perf_counter()
        r = subprocess.run([exe], cwd=d, stdin=subprocess.DEVNULL, capture_output=True, text=True, timeout=300)
        wall = time.perf_counter() - t0
        rec = {"kind": "reference GPU program (unmodified, sm_100)", "wall_seconds": wall, "returncode": r.returncode}
        try:
            row = open(os.path.join(d, "batchTest.csv")).read().strip().splitlines()[-1].split(",")
            rec.update({"keff_printed": float(row[2]), "gpu_seconds": float(row[3]), "cells": int(row[4]), "residual": float(row[5])})
        except Exception as e:
            rec["row_error"] = repr(e) + " | " + r.stdout[-300:]
        img = np.load(os.path.join(ROOT, "tests", "golden", "img_00000.npy"))
        solid = api.structure_from_image(img, 2, 2)
        api.solve2d(solid, p, want_T=False)
        g = api.solve2d(solid, p, want_T=False)
        rec.update({"keffb200_gpu_ms": g["gpu_ms"], "keffb200_iters": g["iters"], "keffb200_keff": g["keff"],
                    "discrete_keff": 4.9003936365, "sample": "Keff2DGPU shipped input: 00000.jpg, MeshAmp 2 (256 x 256), Convergence 1e-5"})
        if "gpu_seconds" in rec and g["gpu_ms"] > 0:
            rec["speedup_device_time"] = rec["gpu_seconds"] * 1e3 / g["gpu_ms"]
        return rec
    except Exception as e:
        return {"error": repr(e)}
    finally:
        shutil.rmtree(d, ignore_errors=True)


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
                     "frac": achieved / peak, "frac_kind": "HBM-equivalent work rate (SURVEY 8d: 16 B x cell-updates), NOT DRAM "
                     "bandwidth: see hbm_true for the DRAM fraction and issue for the bound that applies",
                     "traffic": ncu_traffic("k_batch2d_mg", n_img), "peak_kind": peak_kind,
                     "algorithmic_bytes_per_launch": 16.0 * updates_per_step,
                     "note": "HBM-equivalent: 16 B x cell-updates / kernel time (4.31 stencil passes per CG iteration, see "
                             "DESIGN 3).  The whole solve of an image runs inside one SM (shared memory, registers, an "
                             "L2-resident slice), so this kernel is issue/latency bound, not HBM bound -- its real DRAM "
                             "traffic is `traffic` -- and the figure is a work rate, not a bandwidth claim (SURVEY 8d, "
                             "config-2 caveat); the HBM-bound kernel of the path is reported under stencil3d",
                     "gcell_updates_per_s": updates_per_step / (k_ms * 1e-3) / 1e9},
    }

    # ---- e2e from PAGEABLE host memory (what a reference-style caller with a std::vector passes) -------------
    pageable = np.array(harr, copy=True)
    api.solve2d_batch(pageable, p)
    dt_p, _ = timed(lambda: api.solve2d_batch(pageable, p))
    out["e2e"]["pageable_value"] = world * n_img * args.steps / dt_p
    out["e2e"]["pageable_note"] = "same call with the structures in ordinary (pageable) host memory"
    del pageable

    # ---- what bounds the batch kernel (SURVEY 8d, config-2 caveat): true DRAM fraction and issue-slot roofline ----
    tr = out["roofline"]["traffic"]
    if tr:
        out["roofline"]["hbm_true"] = {"achieved": tr / (k_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                       "frac": tr / (k_ms * 1e-3) / 1e9 / peak,
                                       "note": "measured DRAM bytes (ncu) / kernel time / measured HBM peak"}
    try:
        wi = float(json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["k_batch2d_mg"]["warp_inst_per_unit"])
        sm_hz = (clocks.get("sm_mhz") or 1965.0) * 1e6
        sms = torch.cuda.get_device_properties(local).multi_processor_count
        issue_peak = sms * 4 * sm_hz                                   # 4 schedulers per SM, one warp instruction per clock each
        issue = wi * n_img / (k_ms * 1e-3)
        out["roofline"]["issue"] = {"bound": "issue", "achieved": issue / 1e12, "peak": issue_peak / 1e12,
                                    "unit": "T warp-inst/s", "frac": issue / issue_peak,
                                    "note": "warp instructions per image (ncu smsp__inst_executed.sum of the committed capture) "
                                            "x images / kernel time, against SMs x 4 schedulers x SM clock under load"}
    except Exception:
        pass

    # ---- configs[1] AS NAMED at N > 1: the same 10,000 images split across the GPUs (strong scaling) -----------
    if world > 1:
        i0, cnt = KD_image_shard(n_img, rank, world)
        shard_dev = devm[i0:i0 + cnt].contiguous()
        shard_host = host[i0:i0 + cnt].contiguous().pin_memory().numpy()
        api.solve2d_batch(shard_dev, p)
        dt_s, _ = timed(lambda: api.solve2d_batch(shard_dev, p))
        api.solve2d_batch(shard_host, p)
        dt_se, _ = timed(lambda: api.solve2d_batch(shard_host, p))
        out["batch_split"] = {"workload": f"configs[1] as named: {n_img} images split over {world} GPUs ({cnt} on rank {rank}), "
                              "no data-path collective", "scaling": "strong", "value": n_img * args.steps / dt_s,
                              "e2e": n_img * args.steps / dt_se, "unit": UNIT, "ms_per_step": 1e3 * dt_s / args.steps,
                              "e2e_ms_per_step": 1e3 * dt_se / args.steps}
        del shard_dev

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
        try:                                                        # the same kernel as the multigrid iteration launches it
            pf = api.default_params(flags=32)                         # KEFFB200_F_APPLY_F32: p stored in FP32
            msf = api.bench_apply(vol, pf, reps=20, device=local)
            gbf = 12.0 * n3 ** 3 / (msf * 1e-3) / 1e9
            out["stencil3d_f32in"] = {"workload": "configs[3], input field in FP32 (p of the multigrid-preconditioned iteration)",
                                      "kernel": "k_stencil_tma<APPLY, float>", "ms_per_sweep": msf,
                                      "gcell_updates_per_s": n3 ** 3 / (msf * 1e-3) / 1e9, "bound": "hbm", "achieved": gbf,
                                      "peak": peak, "unit": "GB/s", "frac": gbf / peak,
                                      "algorithmic_bytes_per_launch": 12.0 * n3 ** 3,
                                      "traffic": ncu_traffic("k_stencil_tma_apply_f32in", n3 ** 3)}
        except Exception as e:
            out["stencil3d_f32in"] = {"error": repr(e)}
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

    if world == 1 and not args.no_3d and torch.cuda.get_device_properties(local).total_memory > 100e9:
        # configs[4] on ONE GPU: the 1024^3 volume fits in 180 GB (about 60 GB of workspace)
        del devm
        torch.cuda.empty_cache()
        n4 = 1024
        vol = api.bernoulli_volume_dev((n4, n4, n4), 0.5, 1234, device=local)
        api.solve3d(vol, p, want_T=False, device=local)
        r4 = api.solve3d(vol, p, want_T=False, device=local)
        out["solve3d_1024"] = {"workload": "configs[4] on one GPU: 1024^3 random binary volume, ks:kf 10:1",
                               "gpu_ms": r4["gpu_ms"], "iters": r4["iters"], "ms_per_iter": r4["gpu_ms"] / max(r4["iters"], 1),
                               "keff": r4["keff"], "rel_residual": r4["rel_residual"], "status": r4["status"]}
        del vol
        K.shutdown()                                                 # give the 60 GB back before the CPU legs
        K.init(local)

    if world > 1 and not args.no_3d:
        from keff_cfd_b200 import dist as KD
        try:
            del devm
        except NameError:
            pass
        torch.cuda.empty_cache()
        KD.init_comm(rank, world, local)
        ps = api.default_params()

        def slab_case(nxs, nzg, want_T=False):
            z0, nzl = KD.slab_extent(nzg, rank, world)
            loc = api.bernoulli_volume_dev((nzl, nxs, nxs), 0.5, 1234, z0=z0, device=local)
            halo = KD.exchange_structure_halos(loc, rank, world)
            del loc
            for _ in range(2):                                      # first call allocates the workspace
                barrier()
                r = KD.solve3d_slab(halo, z0, nzg, ps, want_T=want_T)
            r["gpu_ms"] = max_over_ranks(float(r["gpu_ms"]))
            return r, z0, nzl

        def slab_record(r, nxs, nzg, what):
            return {"workload": what, "gpu_ms": r["gpu_ms"], "solves_per_s": 1e3 / r["gpu_ms"], "iters": r["iters"],
                    "ms_per_iter": r["gpu_ms"] / max(r["iters"], 1), "keff": r["keff"], "rel_residual": r["rel_residual"],
                    "status": r["status"], "gcell_iters_per_s": nxs * nxs * nzg * r["iters"] / (r["gpu_ms"] * 1e-3) / 1e9,
                    "transport": "peer memory (NVLink stores from the solver's kernels)" if KD.peer_memory() else "NCCL"}

        # (1) slab-vs-single parity on the record: a 256^3 splitmix volume in `world` slabs against rank 0 alone
        npar = 256
        rp, z0p, nzlp = slab_case(npar, npar, want_T=True)
        # Device-wide sync before the field leaves: without it the 8-GPU run of this leg twice delivered rank 0's slab as
        # stale memory (Keff, computed inside the solve, was right; the other seven slabs too), while the stand-alone
        # replay of the same sequence (scratch/dbg_slabT.py) never did -- an ordering hazard between the NCCL stream of
        # the gather and the generator / allocator traffic of this script, not of the solve.  The sums make it checkable.
        torch.cuda.synchronize()
        sum_before = float(rp["T"].sum().item())                    # each rank's own field before it travels
        parts = [torch.empty_like(rp["T"]) for _ in range(world)]
        dist.all_gather(parts, rp["T"])
        torch.cuda.synchronize()
        sums = [None] * world
        dist.all_gather_object(sums, (sum_before, float(rp["T"].sum().item()), float(parts[rank].sum().item())))
        if rank == 0:
            full = api.bernoulli_volume_dev((npar, npar, npar), 0.5, 1234, device=local)
            s1 = api.solve3d(full, ps, device=local)
            torch.cuda.synchronize()
            per_slab = [float((parts[q] - s1["T"][q * nzlp:(q + 1) * nzlp]).abs().max().item()) for q in range(world)]
            out["slab_parity"] = {"workload": f"256^3 random binary volume: {world} z-slabs vs the same solve on one GPU",
                                  "keff_slab": rp["keff"], "keff_single": s1["keff"],
                                  "keff_rel_diff": abs(rp["keff"] - s1["keff"]) / abs(s1["keff"]),
                                  "T_maxdiff": max(per_slab), "T_maxdiff_per_slab": per_slab,
                                  "rel_residual_slab": rp["rel_residual"], "rel_residual_single": s1["rel_residual"],
                                  "iters_slab": rp["iters"], "iters_single": s1["iters"], "status": rp["status"]}
            if max(per_slab) > 1e-6:                                # diagnostics for the record: which field is off, and how
                out["slab_parity"]["diagnostic"] = {
                    "own_field_sums_before_gather_after_gather_own_part": sums,
                    "single_T_range": [float(s1["T"].min().item()), float(s1["T"].max().item())],
                    "single_T_zeros": int((s1["T"] == 0).sum().item()),
                    "slab_T_ranges": [[float(q.min().item()), float(q.max().item())] for q in parts],
                    "slab_T_zeros": [int((q == 0).sum().item()) for q in parts]}
            del full, s1
        del parts, rp
        torch.cuda.empty_cache()
        barrier()
        # (2) weak scaling: 1024 x 1024 x 128 planes per GPU (134 M cells per GPU = the 512^3 single-GPU case)
        nxs = 1024
        rw, _, _ = slab_case(nxs, 128 * world)
        if rank == 0:
            out["slab3d"] = slab_record(rw, nxs, 128 * world, f"1024x1024x{128 * world} random binary volume in {world} z-slabs of "
                                        "128 planes (configs[4] at 8 GPUs), ks:kf 10:1")
            out["slab3d"]["scaling"] = "weak (compare gpu_ms with solve3d.gpu_ms of the same line: 512^3, same cells per GPU)"
            if "solve3d" in out:
                out["slab3d"]["weak_efficiency"] = out["solve3d"]["gpu_ms"] / rw["gpu_ms"]
        # (3) configs[4] AS NAMED: the 1024^3 volume in `world` slabs (strong scaling)
        if world != 8:
            rs, _, _ = slab_case(nxs, nxs)
        else:
            rs = rw
        if rank == 0:
            out["slab3d_strong"] = slab_record(rs, nxs, nxs, f"configs[4] as named: 1024^3 random binary volume in {world} z-slabs "
                                               f"of {nxs // world} planes, ks:kf 10:1")
            out["slab3d_strong"]["scaling"] = "strong (1-GPU time of the same volume: solve3d_1024 of the N = 1 line)"

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
        if world == 1:
            out["cpu_baseline_others"] = cpu_other_configs(refbind, api, p)
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
