"""Golden vectors for the host programs' image reader (keff_cfd_b200/csrc/host/image_gray.hpp): small files in every
format it reads, each decoded by the REFERENCE's decoder (stb_image.h as vendored by Keff2D, through oracle/_ref/stbgray,
built from /root/reference by oracle/Makefile) -> tests/golden/images/<file> + <file>.raw ("w h channels\\n" + bytes).
Run in the build container (needs /root/reference and PIL):  python oracle/make_image_golden.py"""
import os
import struct
import subprocess
import sys
import zlib

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden", "images")
STB = os.path.join(ROOT, "oracle", "_ref", "stbgray")


def structure(w, h, seed):
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w]
    a = np.zeros((h, w), np.uint8)
    for _ in range(9):
        cx, cy, r = rng.uniform(0, w), rng.uniform(0, h), rng.uniform(4, 14)
        a[(x - cx) ** 2 + (y - cy) ** 2 < r * r] = 255
    return np.clip(a.astype(int) + rng.integers(-35, 35, (h, w)), 0, 255).astype(np.uint8)


def chunk(t, b):
    return struct.pack(">I", len(b)) + t + b + struct.pack(">I", zlib.crc32(t + b) & 0xffffffff)


def png_bytes(w, h, depth, ctype, raw, interlace):
    return (b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, ctype, 0, 0, interlace))
            + chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))


def adam7(img, depth):
    xs, ys, dx, dy = [0, 4, 0, 2, 0, 1, 0], [0, 0, 4, 0, 2, 0, 1], [8, 8, 4, 4, 2, 2, 1], [8, 8, 8, 4, 4, 2, 2]
    raw = b""
    for p in range(7):
        sub = img[ys[p]::dy[p], xs[p]::dx[p]]
        for row in sub:
            raw += b"\x00" + (row.astype(">u2").tobytes() if depth == 16 else row.astype(np.uint8).tobytes())
    return raw


def main():
    if not os.path.exists(STB):
        sys.exit("oracle/_ref/stbgray is missing: run `make -C oracle ref` where /root/reference exists")
    os.makedirs(OUT, exist_ok=True)
    g = structure(77, 53, 1)
    rgb = np.stack([g, structure(77, 53, 2), structure(77, 53, 3)], -1)
    Image.fromarray(g).save(os.path.join(OUT, "gray_base.jpg"), quality=90)
    Image.fromarray(g).save(os.path.join(OUT, "gray_prog.jpg"), quality=85, progressive=True)
    Image.fromarray(g).save(os.path.join(OUT, "gray_prog_lowq.jpg"), quality=35, progressive=True, optimize=True)
    Image.fromarray(rgb).save(os.path.join(OUT, "ycc_prog_420.jpg"), quality=80, progressive=True, subsampling=2)
    Image.fromarray(rgb).save(os.path.join(OUT, "ycc_base_444.jpg"), quality=92, subsampling=0)
    Image.fromarray(g).save(os.path.join(OUT, "gray_restart.jpg"), quality=90, restart_marker_blocks=3)
    Image.fromarray(g).save(os.path.join(OUT, "gray8.png"))
    Image.fromarray(rgb).save(os.path.join(OUT, "rgb8.png"))
    Image.fromarray(np.dstack([rgb, g])).save(os.path.join(OUT, "rgba8.png"))
    Image.fromarray(rgb).convert("P", palette=Image.ADAPTIVE, colors=60).save(os.path.join(OUT, "palette.png"))
    Image.fromarray(g > 150).save(os.path.join(OUT, "bilevel.png"))
    open(os.path.join(OUT, "gray8_adam7.png"), "wb").write(png_bytes(77, 53, 8, 0, adam7(g, 8), 1))
    r16 = rgb.astype(np.uint16) * 251 + 7
    open(os.path.join(OUT, "rgb16_adam7.png"), "wb").write(png_bytes(77, 53, 16, 2, adam7_rgb16(r16), 1))
    Image.fromarray(rgb).save(os.path.join(OUT, "rgb24.bmp"))
    Image.fromarray(g).save(os.path.join(OUT, "palette8.bmp"))
    Image.fromarray(g).save(os.path.join(OUT, "gray.pgm"))
    import shutil
    for f in ("00009.jpg", "00000.jpg"):                      # the reference's shipped structures
        shutil.copy(os.path.join(ROOT, "tests", "golden", "inputs", f), os.path.join(OUT, "shipped_" + f))
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".raw"):
            continue
        subprocess.run([STB, os.path.join(OUT, f), os.path.join(OUT, f + ".raw")], check=True)
        print(f, open(os.path.join(OUT, f + ".raw"), "rb").readline().decode().strip())


def adam7_rgb16(img):
    xs, ys, dx, dy = [0, 4, 0, 2, 0, 1, 0], [0, 0, 4, 0, 2, 0, 1], [8, 8, 4, 4, 2, 2, 1], [8, 8, 8, 4, 4, 2, 2]
    raw = b""
    for p in range(7):
        for row in img[ys[p]::dy[p], xs[p]::dx[p]]:
            raw += b"\x00" + row.reshape(-1).astype(">u2").tobytes()
    return raw


if __name__ == "__main__":
    main()
