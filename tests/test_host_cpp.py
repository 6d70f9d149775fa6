"""The C++ drop-in shell (keff_cfd_b200/csrc/host): input.txt keys, JPG / CSV readers, output rows.
CPU part: decoder and parsers (--check-inputs, no GPU).  GPU part: the programs end to end."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from conftest import GOLD, ROOT, gold_json

HOST = os.path.join(ROOT, "keff_cfd_b200", "csrc", "host")
INPUT_2D = """Input File:
ks: 10
kf: 1
MeshAmpX: {amp}
MeshAmpY: {amp}
InputName: {name}
TR: 1
TL: 0
OutputName: batchTest.csv
printTMap: {tmap}
TMapName: TMAP_00001.csv
printQMap: {tmap}
QMapName: QMAP_00001.csv
Convergence: 0.00001
MaxIter: 500000
Verbose: 1
NumCores: 8
MeshFlag: 0
MaxMesh: 10
RunBatch: {batch}
NumImages: {nimg}
"""


@pytest.fixture(scope="module")
def built():
    subprocess.run(["make", "-C", os.path.join(ROOT, "keff_cfd_b200", "csrc"), "all"], check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return HOST


def _decode(built, path, tmp):
    out = os.path.join(tmp, "out.raw")
    r = subprocess.run([os.path.join(built, "jpeg2raw"), path, out], capture_output=True, text=True)
    if r.returncode:
        return None, r.stderr
    hdr, raw = open(out, "rb").read().split(b"\n", 1)
    w, h, c = map(int, hdr.split())
    return np.frombuffer(raw, np.uint8).reshape(h, w), c


def test_jpeg_decoder_matches_libjpeg(built, tmp_path):
    from PIL import Image
    from keff_cfd_b200 import structures as S
    for f in ("00009.jpg", "00000.jpg"):     # the reference's shipped structures
        a, c = _decode(built, os.path.join(GOLD, "inputs", f), str(tmp_path))
        b = np.asarray(Image.open(os.path.join(GOLD, "inputs", f)))
        assert c == 1 and np.abs(a.astype(int) - b).max() <= 1
        assert np.array_equal(a >= 150, np.load(os.path.join(GOLD, "img_" + f[:5] + ".npy")) >= 150)   # the goldens' structure
    img = S.to_gray(S.disc_image(5, n=100))[:77, :]   # ragged size: partial MCUs
    cases = [("L", dict(quality=95)), ("L", dict(quality=75, optimize=True)), ("RGB", dict(quality=90)),
             ("RGB", dict(quality=90, subsampling=0)), ("L", dict(quality=95, restart_marker_blocks=4))]
    for mode, kw in cases:
        p = str(tmp_path / "t.jpg")
        Image.fromarray(img).convert(mode).save(p, **kw)
        a, c = _decode(built, p, str(tmp_path))
        ref = np.asarray(Image.open(p).convert("L"))
        assert a.shape == ref.shape and c == (1 if mode == "L" else 3)
        assert np.abs(a.astype(int) - ref).max() <= 2 and np.array_equal(a >= 150, ref >= 150)
    p = str(tmp_path / "p.jpg")
    Image.fromarray(img).save(p, progressive=True)
    a, c = _decode(built, p, str(tmp_path))
    ref = np.asarray(Image.open(p).convert("L"))
    assert a is not None and np.abs(a.astype(int) - ref).max() <= 2
    open(p, "wb").write(b"not a jpeg")
    a, msg = _decode(built, p, str(tmp_path))
    assert a is None and "unrecognised image format" in msg
    open(p, "wb").write(b"\xff\xd8\xff\xc3\x00\x0b\x08\x00\x10\x00\x10\x01\x01\x11\x00")   # lossless JPEG frame
    a, msg = _decode(built, p, str(tmp_path))
    assert a is None and "not supported" in msg


def _golden_images():
    d = os.path.join(GOLD, "images")
    return sorted(f for f in os.listdir(d) if not f.endswith(".raw"))


@pytest.mark.parametrize("name", _golden_images())
def test_image_reader_matches_reference_decoder_goldens(built, tmp_path, name):
    """Every format the reader accepts, against what the reference's decoder (stbi_load(..., 1), Keff2D/keff2D.h:300)
    returns for the same file -- bit for bit: a grey level off by one can flip a pixel across the 150 threshold
    (fixtures: oracle/make_image_golden.py)."""
    path = os.path.join(GOLD, "images", name)
    a, c = _decode(built, path, str(tmp_path))
    assert a is not None, c
    hdr, raw = open(path + ".raw", "rb").read().split(b"\n", 1)
    w, h, ch = map(int, hdr.split())
    assert a.shape == (h, w) and c == ch
    assert np.array_equal(a, np.frombuffer(raw, np.uint8).reshape(h, w))


def test_image_reader_matches_reference_decoder_live(built, tmp_path):
    """Freshly generated files (JPEG quality / progressive / chroma-subsampling sweep, PNG and BMP variants) through the
    reference's decoder itself where oracle/_ref/stbgray exists (built from /root/reference by oracle/Makefile)."""
    from PIL import Image
    stb = os.path.join(ROOT, "oracle", "_ref", "stbgray")
    if not os.path.exists(stb):
        pytest.skip("oracle/_ref/stbgray not built (no /root/reference here)")
    rng = np.random.default_rng(11)
    y, x = np.mgrid[0:90, 0:101]
    base = np.zeros((90, 101), np.uint8)
    for _ in range(10):
        cx, cy, r = rng.uniform(0, 101), rng.uniform(0, 90), rng.uniform(5, 18)
        base[(x - cx) ** 2 + (y - cy) ** 2 < r * r] = 255
    g = np.clip(base.astype(int) + rng.integers(-40, 40, base.shape), 0, 255).astype(np.uint8)
    rgb = np.stack([g, g[::-1], rng.integers(0, 256, g.shape, dtype=np.uint8)], -1)
    files = []
    for q in (25, 60, 95):
        for prog in (False, True):
            for arr, tag, extra in ((g, "g", {}), (rgb, "c0", {"subsampling": 0}), (rgb, "c2", {"subsampling": 2})):
                f = str(tmp_path / f"{tag}_{q}_{int(prog)}.jpg")
                Image.fromarray(arr).save(f, quality=q, progressive=prog, optimize=prog, **extra)
                files.append(f)
    for arr, name in ((g, "g.png"), (rgb, "rgb.png"), (np.dstack([rgb, g]), "rgba.png"), (g > 128, "b.png"),
                      (rgb, "rgb.bmp"), (g, "g.bmp"), (g, "g.pgm"), (rgb, "c.ppm")):
        f = str(tmp_path / name)
        Image.fromarray(arr).save(f)
        files.append(f)
    f = str(tmp_path / "pal.png")
    Image.fromarray(rgb).convert("P", palette=Image.ADAPTIVE, colors=33).save(f)
    files.append(f)
    for f in files:
        a, c = _decode(built, f, str(tmp_path))
        out = f + ".stb"
        subprocess.run([stb, f, out], check=True)
        hdr, raw = open(out, "rb").read().split(b"\n", 1)
        w, h, ch = map(int, hdr.split())
        assert a is not None and a.shape == (h, w) and c == ch, f
        assert np.array_equal(a, np.frombuffer(raw, np.uint8).reshape(h, w)), f


def test_image_reader_survives_damaged_files(built, tmp_path):
    """Truncated and bit-flipped files of every format: the reader either decodes or reports an error -- it never
    crashes and never reads outside the file (the tool exits 0 or 1, nothing else)."""
    rng = np.random.default_rng(3)
    tool = os.path.join(built, "jpeg2raw")
    d = os.path.join(GOLD, "images")
    n = 0
    for name in _golden_images():
        raw = open(os.path.join(d, name), "rb").read()
        variants = [raw[:k] for k in sorted(set(int(v) for v in rng.integers(1, len(raw), 6)))]
        for _ in range(6):
            b = bytearray(raw)
            for pos in rng.integers(0, len(raw), 4):
                b[int(pos)] ^= 1 << int(rng.integers(0, 8))
            variants.append(bytes(b))
        for v in variants:
            f = str(tmp_path / ("x" + os.path.splitext(name)[1]))
            open(f, "wb").write(v)
            r = subprocess.run([tool, f, str(tmp_path / "o.raw")], capture_output=True, timeout=20)
            assert r.returncode in (0, 1), (name, r.returncode, r.stderr[-200:])
            n += 1
    assert n > 150


def test_programs_parse_reference_inputs(built, tmp_path):
    d = str(tmp_path)
    shutil.copy(os.path.join(GOLD, "inputs", "00009.jpg"), d)
    open(os.path.join(d, "input.txt"), "w").write(INPUT_2D.format(amp=2, name="00009.jpg", tmap=0, batch=0, nimg=1))
    r = subprocess.run([os.path.join(built, "keff2D"), "--check-inputs"], cwd=d, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    # reference porosity of 00009.jpg = 0.3736572266 (its CSV row), 256 x 256 cells after MeshAmp 2
    assert "128x128 porosity 0.3736572266 cells 65536 ks 10 kf 1 TL 0 TR 1 tol 1e-05 maxit 500000" in r.stdout
    xyz = np.array([[1, 2, 3], [0, 0, 0], [3, 1, 2]])
    np.savetxt(os.path.join(d, "s.csv"), xyz, fmt="%d", delimiter=",")
    open(os.path.join(d, "in3.txt"), "w").write(
        "Input File:\nks: 10\nkf: 1\nWidth: 4\nHeight: 4\nDepth: 4\nInputName: s.csv\nTR: 1\nTL: 0\nOutputName: o.csv\n")
    r = subprocess.run([os.path.join(built, "keff3D"), "in3.txt", "--check-inputs"], cwd=d, capture_output=True, text=True)
    assert r.returncode == 0 and "4x4x4 solid voxels 3" in r.stdout
    r = subprocess.run([os.path.join(built, "keff2D"), "missing.txt"], cwd=d, capture_output=True, text=True)
    assert r.returncode == 1 and "cannot open" in r.stderr


def test_keff2d_program_accepts_the_formats_the_reference_reads(built, tmp_path):
    """InputName may be a progressive JPEG, a PNG or a BMP (stb_image reads them, Keff2D/keff2D.h:300): the drop-in
    program thresholds the same bytes -- checked through the porosity it prints for a structure saved losslessly."""
    from PIL import Image
    from keff_cfd_b200 import structures as S
    d = str(tmp_path)
    img = S.to_gray(S.disc_image(7, n=96))[:80, :]
    fluid = float((img < 150).mean())
    for name in ("s.png", "s.pgm"):
        Image.fromarray(img).save(os.path.join(d, name))
        open(os.path.join(d, "input.txt"), "w").write(INPUT_2D.format(amp=1, name=name, tmap=0, batch=0, nimg=1))
        r = subprocess.run([os.path.join(built, "keff2D"), "--check-inputs"], cwd=d, capture_output=True, text=True)
        assert r.returncode == 0, (name, r.stderr)
        assert "96x80 porosity %.10f cells 7680" % fluid in r.stdout, (name, r.stdout)
    # a palettised BMP decodes to 3 channels in the reference's decoder: single-image mode refuses it exactly as the
    # reference does (Keff2D/keff2D.cpp:60-63)
    Image.fromarray(img).save(os.path.join(d, "s.bmp"))
    open(os.path.join(d, "input.txt"), "w").write(INPUT_2D.format(amp=1, name="s.bmp", tmap=0, batch=0, nimg=1))
    r = subprocess.run([os.path.join(built, "keff2D"), "--check-inputs"], cwd=d, capture_output=True, text=True)
    assert r.returncode == 1 and "Current number of channels = 3" in r.stdout
    Image.fromarray(img).save(os.path.join(d, "p.jpg"), quality=95, progressive=True)   # lossy: against the decoded bytes
    a, c = _decode(built, os.path.join(d, "p.jpg"), d)
    open(os.path.join(d, "input.txt"), "w").write(INPUT_2D.format(amp=1, name="p.jpg", tmap=0, batch=0, nimg=1))
    r = subprocess.run([os.path.join(built, "keff2D"), "--check-inputs"], cwd=d, capture_output=True, text=True)
    assert r.returncode == 0 and "96x80 porosity %.10f cells 7680" % float((a < 150).mean()) in r.stdout, r.stdout + r.stderr
    open(os.path.join(d, "bad.gif"), "wb").write(b"GIF89a" + bytes(40))
    open(os.path.join(d, "input.txt"), "w").write(INPUT_2D.format(amp=1, name="bad.gif", tmap=0, batch=0, nimg=1))
    r = subprocess.run([os.path.join(built, "keff2D"), "--check-inputs"], cwd=d, capture_output=True, text=True)
    assert r.returncode == 1 and "unrecognised image format" in r.stderr


@pytest.mark.gpu
def test_keff2d_program_end_to_end(built, tmp_path):
    d = str(tmp_path)
    shutil.copy(os.path.join(GOLD, "inputs", "00009.jpg"), d)
    open(os.path.join(d, "input.txt"), "w").write(INPUT_2D.format(amp=1, name="00009.jpg", tmap=1, batch=0, nimg=1))
    r = subprocess.run([os.path.join(built, "keff2D")], cwd=d, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr + r.stdout
    row = open(os.path.join(d, "batchTest.csv")).read().strip().split(",")
    ref_row = gold_json("ref2d_00009.json")["binary_row_shipped_keys_1core"].split(",")
    assert len(row) == len(ref_row) == 16                                  # same columns as the reference row
    want = gold_json("discrete.json")["img_00009_amp1"]["keff"]
    assert abs(float(row[0]) - want) / want < 1e-5
    assert abs(float(row[1]) - float(row[2])) < 1e-4                       # QL == QR at convergence
    assert row[4:10] == ref_row[4:10] and row[10:14] == ref_row[10:14]     # tol,name,nElements,amps,porosity,ks,kf,tl,tr
    tm = open(os.path.join(d, "TMAP_00001.csv")).read().splitlines()
    qm = open(os.path.join(d, "QMAP_00001.csv")).read().splitlines()
    assert tm[0] == "x,y,T" and len(tm) == 1 + 128 * 128 and qm[0] == "x,y,Q" and len(qm) == 1 + 128 * 129
    # x-flux through every vertical line of faces is the same at steady state (energy conservation)
    q = np.array([float(l.split(",")[2]) for l in qm[1:]]).reshape(128, 129)
    assert np.abs(q.sum(axis=0) - q.sum(axis=0)[0]).max() < 1e-3


@pytest.mark.gpu
def test_keff2d_batch_and_keff3d_programs(built, tmp_path, R):
    from PIL import Image
    from keff_cfd_b200 import structures as S
    d = str(tmp_path)
    for i in range(3):
        Image.fromarray(S.to_gray(S.disc_image(i, n=64))).save(os.path.join(d, "%05d.jpg" % i), quality=95)
    open(os.path.join(d, "input.txt"), "w").write(INPUT_2D.format(amp=1, name="x.jpg", tmap=0, batch=1, nimg=3))
    r = subprocess.run([os.path.join(built, "keff2D")], cwd=d, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    rows = [l.split(",") for l in open(os.path.join(d, "batchTest.csv")).read().strip().splitlines()]
    assert [x[5] for x in rows] == ["00000.csv", "00001.csv", "00002.csv"]   # reference naming (keff2D.h:271)
    for i, row in enumerate(rows):
        want = R.orc_discrete(S.to_gray(S.disc_image(i, n=64)))["keff"]
        assert abs(float(row[0]) - want) / want < 1e-5
    m = S.bernoulli_volume((12, 12, 12), 0.4, 3)
    z, y, x = np.nonzero(m)
    np.savetxt(os.path.join(d, "vol.csv"), np.stack([x, y, z], 1), fmt="%d", delimiter=",")
    open(os.path.join(d, "in3.txt"), "w").write(
        "Input File:\nks: 10\nkf: 1\nWidth: 12\nHeight: 12\nDepth: 12\nMeshAmpX: 1\nMeshAmpY: 1\nMeshAmpZ: 1\n"
        "InputName: vol.csv\nTR: 1\nTL: 0\nOutputName: out3.csv\nprintTMap: 1\nTMapName: T3.csv\nprintQMap: 0\n"
        "Convergence: 0.00001\nMaxIter: 500000\nVerbose: 0\nNumCores: 1\nMeshFlag: 0\nRunBatch: 0\nNumImages: 1\n")
    r = subprocess.run([os.path.join(built, "keff3D"), "in3.txt"], cwd=d, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    row = open(os.path.join(d, "out3.csv")).read().strip().split(",")
    want = R.orc_discrete(m)["keff"]
    assert len(row) == 17 and abs(float(row[0]) - want) / want < 1e-5 and row[5] == "vol.csv" and row[6] == "1728"
    t = open(os.path.join(d, "T3.csv")).read().splitlines()
    assert t[0] == "x,y,z,T" and len(t) == 1 + 1728


def test_batch_decode_pipeline_on_cpu(built, tmp_path):
    """Batch mode decodes %05d.jpg on NumCores threads in chunks; --check-inputs runs
    that whole pipeline without a GPU.  A missing image must fail the run with its name."""
    from PIL import Image
    from keff_cfd_b200 import structures as S
    d = str(tmp_path)
    for i in range(37):
        Image.fromarray(S.to_gray(S.disc_image(i, n=32))).save(os.path.join(d, "%05d.jpg" % i), quality=95)
    open(os.path.join(d, "input.txt"), "w").write(INPUT_2D.format(amp=2, name="x.jpg", tmap=0, batch=1, nimg=37))
    r = subprocess.run([os.path.join(built, "keff2D"), "input.txt", "--check-inputs"], cwd=d, capture_output=True, text=True)
    assert r.returncode == 0 and "batch ok: 37 images 64x64" in r.stdout, (r.stdout, r.stderr)
    os.remove(os.path.join(d, "00020.jpg"))
    r = subprocess.run([os.path.join(built, "keff2D"), "input.txt", "--check-inputs"], cwd=d, capture_output=True, text=True)
    assert r.returncode == 1 and "00020.jpg" in r.stderr


@pytest.mark.gpu
def test_mesh_convergence_modes_against_oracle(built, tmp_path, R):
    """MeshFlag: 1 (Keff2D/keff2D.cpp:193-332, Keff3D/keff3D.cpp:187-331): one row per mesh m = 1..MaxMesh, each
    against the discrete solution of the m-times amplified structure; warm-started meshes take fewer iterations."""
    from PIL import Image
    from keff_cfd_b200 import structures as S
    from keff_cfd_b200.api import structure_from_image
    d = str(tmp_path)
    img = S.to_gray(S.disc_image(11, n=32))
    Image.fromarray(img).save(os.path.join(d, "m.jpg"), quality=95)
    txt = INPUT_2D.format(amp=1, name="m.jpg", tmap=0, batch=0, nimg=1).replace("MeshFlag: 0", "MeshFlag: 1").replace("MaxMesh: 10", "MaxMesh: 3")
    open(os.path.join(d, "input.txt"), "w").write(txt)
    r = subprocess.run([os.path.join(built, "keff2D")], cwd=d, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr + r.stdout
    rows = [l.split(",") for l in open(os.path.join(d, "batchTest.csv")).read().strip().splitlines()]
    assert len(rows) == 3 and [int(x[6]) for x in rows] == [32 * 32, 64 * 64, 96 * 96]
    for m, row in zip((1, 2, 3), rows):
        want = R.orc_discrete(S.to_gray(structure_from_image(img, m, m)))["keff"]
        assert abs(float(row[0]) - want) / want < 1e-5, (m, row[0], want)
    assert int(rows[1][3]) <= int(rows[0][3]) and int(rows[2][3]) <= int(rows[0][3])      # warm starts help
    # 3D: a small volume, m = 1, 2; plus the Q-map the reference never wrote
    vol = S.bernoulli_volume((6, 8, 10), 0.4, 3)
    z, y, x = np.nonzero(vol)
    np.savetxt(os.path.join(d, "vol.csv"), np.stack([x, y, z], 1), fmt="%d", delimiter=",")
    base = ("Input File:\nks: 10\nkf: 1\nWidth: 10\nHeight: 8\nDepth: 6\nMeshAmpX: 1\nMeshAmpY: 1\nMeshAmpZ: 1\n"
            "InputName: vol.csv\nTR: 1\nTL: 0\nOutputName: out3.csv\nprintTMap: 0\nTMapName: T3.csv\nprintQMap: {q}\nQMapName: Q3.csv\n"
            "Convergence: 0.00001\nMaxIter: 500000\nVerbose: 0\nNumCores: 1\nMeshFlag: {mf}\nMaxMesh: 2\nRunBatch: 0\nNumImages: 1\n")
    open(os.path.join(d, "in3.txt"), "w").write(base.format(q=0, mf=1))
    r = subprocess.run([os.path.join(built, "keff3D"), "in3.txt"], cwd=d, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr + r.stdout
    rows = [l.split(",") for l in open(os.path.join(d, "out3.csv")).read().strip().splitlines()]
    assert len(rows) == 2 and [int(x[6]) for x in rows] == [480, 3840]
    for m, row in zip((1, 2), rows):
        amp = np.repeat(np.repeat(np.repeat(vol, m, axis=0), m, axis=1), m, axis=2)
        want = R.orc_discrete(amp)["keff"]
        assert abs(float(row[0]) - want) / want < 1e-5, (m, row[0], want)
    os.remove(os.path.join(d, "out3.csv"))
    open(os.path.join(d, "in3.txt"), "w").write(base.format(q=1, mf=0))
    r = subprocess.run([os.path.join(built, "keff3D"), "in3.txt"], cwd=d, capture_output=True, text=True)
    assert r.returncode == 0 and "not currently available" not in r.stdout, r.stderr + r.stdout
    q = open(os.path.join(d, "Q3.csv")).read().splitlines()
    assert q[0] == "x,y,z,Q" and len(q) == 1 + 6 * 8 * 11
    keff = float(open(os.path.join(d, "out3.csv")).read().split(",")[0])
    flux = np.array([float(l.split(",")[3]) for l in q[1:]]).reshape(6, 8, 11)
    assert np.abs(flux.sum(axis=(0, 1)) - keff).max() < 1e-4                                # every plane carries Keff * (TR - TL)


@pytest.mark.gpu
def test_keff2d_batch_uses_every_gpu(built, tmp_path, R):
    """Batch mode binds every visible GPU (keffb200_init_multi) and shards each chunk over them."""
    import torch
    from PIL import Image
    from keff_cfd_b200 import structures as S
    n_gpu = torch.cuda.device_count()
    d = str(tmp_path)
    n_img = 256
    for i in range(n_img):
        Image.fromarray(S.to_gray(S.disc_image(i, n=32))).save(os.path.join(d, "%05d.jpg" % i), quality=95)
    open(os.path.join(d, "input.txt"), "w").write(INPUT_2D.format(amp=1, name="x.jpg", tmap=0, batch=1, nimg=n_img))
    r = subprocess.run([os.path.join(built, "keff2D")], cwd=d, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert f"{n_gpu} GPU" in r.stdout, r.stdout
    rows = [l.split(",") for l in open(os.path.join(d, "batchTest.csv")).read().strip().splitlines()]
    assert len(rows) == n_img
    for i in (0, 100, 255):                                                                  # first, middle, last shard
        want = R.orc_discrete(S.to_gray(S.disc_image(i, n=32)))["keff"]
        assert abs(float(rows[i][0]) - want) / want < 1e-5
