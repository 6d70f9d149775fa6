#!/usr/bin/env python
"""TEST INFRASTRUCTURE: generate the committed fixtures under tests/golden/ by running the UNMODIFIED
reference (oracle/_ref, built by oracle/Makefile from /root/reference) in THIS container.

    python oracle/make_golden.py prep        # inputs: decoded images / packed Schwarz masks
    python oracle/make_golden.py ref2d       # reference 2D runs (seconds to minutes)
    python oracle/make_golden.py ref3d NAME TOL   # one reference 3D run (8 min @1e-5, ~50 min @1e-10)
    python oracle/make_golden.py binary2d    # the stand-alone reference program on 00009.jpg

Everything written is small (<= 256 KiB per file).  /root/reference is needed only here; tests read
tests/golden/ only.
"""
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refbind  # noqa: E402

REF = "/root/reference"
GOLD = os.path.join(ROOT, "tests", "golden")


def _update_json(name, key, val):
    path = os.path.join(GOLD, name)
    d = json.load(open(path)) if os.path.exists(path) else {}
    d[key] = val
    json.dump(d, open(path, "w"), indent=1, sort_keys=True)


def prep():
    from PIL import Image
    os.makedirs(GOLD, exist_ok=True)
    for src, dst in ((f"{REF}/Keff2D/00009.jpg", "img_00009"), (f"{REF}/Keff2DGPU/00000.jpg", "img_00000")):
        im = Image.open(src)
        assert im.mode == "L", im.mode
        a = np.asarray(im, dtype=np.uint8)
        np.save(os.path.join(GOLD, dst + ".npy"), a)
        print(dst, a.shape, "fluid fraction", float((a < 150).mean()),
              "pixels within 40 of threshold:", int(((a > 110) & (a < 190)).sum()))
    for name in ("Schwarz091", "Schwarz163"):
        xyz = np.loadtxt(f"{REF}/Keff3D/{name}.csv", delimiter=",", dtype=np.int64)
        m = np.zeros((128, 128, 128), dtype=np.uint8)
        m[xyz[:, 2], xyz[:, 1], xyz[:, 0]] = 1  # keff3D.h:269-277
        np.save(os.path.join(GOLD, name.lower() + "_bits.npy"), np.packbits(m))
        print(name, "rows", len(xyz), "solid voxels", int(m.sum()), "solid fraction", float(m.mean()))


def ref2d():
    for tag in ("00009", "00000"):
        img = np.load(os.path.join(GOLD, f"img_{tag}.npy"))
        for amp, tol, keepT in ((1, 1e-5, True), (1, 1e-8, False), (1, 1e-12, True), (2, 1e-5, False)):
            t0 = time.time()
            r = refbind.ref2d_solve(img, amp=(amp, amp), tol=tol)
            dt = time.time() - t0
            key = f"amp{amp}_tol{tol:g}"
            rec = {k: r[k] for k in ("keff", "Q1", "Q2", "iters", "residual", "porosity")}
            rec["seconds"] = round(dt, 2)
            _update_json(f"ref2d_{tag}.json", key, rec)
            if keepT:
                np.save(os.path.join(GOLD, f"ref2d_{tag}_{key}_T.npy"), r["T"])
            print(tag, key, rec, flush=True)
    # ParallelGS with one thread must equal SerialGS bit for bit
    img = np.load(os.path.join(GOLD, "img_00009.npy"))
    a = refbind.ref2d_solve(img, solver="parallel", threads=1)
    b = refbind.ref2d_solve(img, solver="serial")
    assert a["keff"] == b["keff"] and a["iters"] == b["iters"]
    print("ParallelGS(1 thread) == SerialGS:", a["keff"], a["iters"])


def ref3d(name, tol):
    tol = float(tol)
    m = np.unpackbits(np.load(os.path.join(GOLD, name.lower() + "_bits.npy"))).reshape(128, 128, 128)
    t0 = time.time()
    r = refbind.ref3d_solve(m, tol=tol, threads=1)
    dt = time.time() - t0
    rec = {k: r[k] for k in ("keff", "Q1", "Q2", "iters")}
    rec["seconds"] = round(dt, 1)
    _update_json(f"ref3d_{name.lower()}.json", f"tol{tol:g}", rec)
    # temperature on a stride-4 lattice (32^3 doubles = 256 KiB)
    np.save(os.path.join(GOLD, f"ref3d_{name.lower()}_tol{tol:g}_Tsub4.npy"), r["T"][::4, ::4, ::4].copy())
    print(name, tol, rec, flush=True)


def binary2d():
    """The reference's own executable, end to end (stb_image decode, input.txt, CSV row)."""
    d = tempfile.mkdtemp()
    shutil.copy(f"{REF}/Keff2D/00009.jpg", d)
    txt = open(f"{REF}/Keff2D/input.txt").read()
    txt = txt.replace("InputName: 00001.jpg", "InputName: 00009.jpg").replace("NumCores: 8", "NumCores: 1")
    txt = txt.replace("printTMap: 1", "printTMap: 0").replace("printQMap: 1", "printQMap: 0")
    open(os.path.join(d, "input.txt"), "w").write(txt)
    subprocess.run([os.path.join(ROOT, "oracle", "_ref", "keff2D")], cwd=d, check=True, stdout=subprocess.DEVNULL)
    row = open(os.path.join(d, "batchTest.csv")).read().strip()
    _update_json("ref2d_00009.json", "binary_row_shipped_keys_1core", row)
    print(row)
    shutil.rmtree(d)


def discrete():
    """Discrete solutions (oracle PCG on the reference's own assembled system, eps 1e-12) of every
    fixture structure: the limit the tightened reference converges to (checked by
    tests/test_oracle.py against the ref2d_*/ref3d_* records above)."""
    from keff_cfd_b200 import structures as S
    from keff_cfd_b200.api import structure_from_image
    out = {}
    for tag in ("00009", "00000"):
        img = np.load(os.path.join(GOLD, f"img_{tag}.npy"))
        for amp in (1, 2):
            g = S.to_gray(structure_from_image(img, amp, amp))
            d = refbind.orc_discrete(g)
            out[f"img_{tag}_amp{amp}"] = {k: d[k] for k in ("keff", "Q1", "Q2", "iters", "relres")}
            if amp == 1:
                np.save(os.path.join(GOLD, f"discrete_img_{tag}_T.npy"), d["T"])
    for name in ("schwarz163", "schwarz091"):
        m = np.unpackbits(np.load(os.path.join(GOLD, name + "_bits.npy"))).reshape(128, 128, 128)
        d = refbind.orc_discrete(m)
        out[name] = {k: d[k] for k in ("keff", "Q1", "Q2", "iters", "relres")}
        np.save(os.path.join(GOLD, f"discrete_{name}_Tsub4.npy"), d["T"][::4, ::4, ::4].copy())
        # x<->y and x<->z transposes must give the same Keff (the TPMS structures are symmetric)
        out[name]["keff_transposed_xy"] = refbind.orc_discrete(np.ascontiguousarray(m.transpose(0, 2, 1)))["keff"]
    for n in (32, 64):
        m = S.bernoulli_volume((n, n, n), 0.5, 1234)
        d = refbind.orc_discrete(m)
        out[f"bernoulli_{n}_seed1234"] = {k: d[k] for k in ("keff", "Q1", "Q2", "iters", "relres")}
    for i in range(4):
        d = refbind.orc_discrete(S.to_gray(S.disc_image(i)))
        out[f"disc_image_{i}"] = {k: d[k] for k in ("keff", "Q1", "Q2", "iters", "relres")}
    json.dump(out, open(os.path.join(GOLD, "discrete.json"), "w"), indent=1, sort_keys=True)
    for k, v in out.items():
        print(k, v)


def discrete_big():
    """Discrete solutions at the bench's own generators and larger sizes: the splitmix Bernoulli volumes at 128^3
    and 256^3 (configs[3]/[4] scaled to what a CPU finishes: BASELINE.md 4.5) and 256 of the 10,000 bench images."""
    from keff_cfd_b200 import structures as S
    out = {}
    for n, sub in ((128, 4), (256, 8)):
        m = S.bernoulli_volume((n, n, n), 0.5, 1234)
        t0 = time.time()
        d = refbind.orc_discrete(m)
        out[f"bernoulli_{n}_seed1234"] = {k: d[k] for k in ("keff", "Q1", "Q2", "iters", "relres")}
        out[f"bernoulli_{n}_seed1234"]["seconds"] = round(time.time() - t0, 1)
        np.save(os.path.join(GOLD, f"discrete_bernoulli_{n}_Tsub{sub}.npy"), d["T"][::sub, ::sub, ::sub].copy())
        print(n, out[f"bernoulli_{n}_seed1234"], flush=True)
    idx = [39 * i for i in range(256)]
    keff = []
    for i in idx:
        d = refbind.orc_discrete(S.to_gray(S.disc_image(i)))
        keff.append([i, d["keff"], d["Q1"], d["Q2"], d["relres"]])
    np.save(os.path.join(GOLD, "discrete_bench_images.npy"), np.array(keff))
    json.dump(out, open(os.path.join(GOLD, "discrete_big.json"), "w"), indent=1, sort_keys=True)
    print("bench images:", len(keff), "max relres", max(k[4] for k in keff))


if __name__ == "__main__":
    cmd = sys.argv[1]
    {"prep": prep, "ref2d": ref2d, "binary2d": binary2d, "discrete": discrete, "discrete_big": discrete_big}.get(
        cmd, lambda: ref3d(*sys.argv[2:4]))()
