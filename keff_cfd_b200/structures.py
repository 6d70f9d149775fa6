"""Reproducible synthetic structures.  Counter-based: value(i) = splitmix64(seed ^ i),
so the same structure can be produced from Python, C++ or CUDA (keffb200_gen_bernoulli_dev)."""
import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x):
    """splitmix64 finaliser on a uint64 array (wrap-around arithmetic)."""
    with np.errstate(over="ignore"):
        z = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def bernoulli_threshold(p):
    return np.uint64(min(int(p * 18446744073709551616.0), 0xFFFFFFFFFFFFFFFF))


def bernoulli_volume(shape, p=0.5, seed=1234, z0=0):
    """iid Bernoulli(p) solid voxels, u8 [nz, ny, nx]; voxel i (global linear index, z offset z0
    for slabs) is solid iff splitmix64(seed ^ i) < p * 2^64."""
    nz, ny, nx = shape
    thr = bernoulli_threshold(p)
    out = np.empty(shape, dtype=np.uint8)
    plane = ny * nx
    for z in range(nz):
        i = np.arange(plane, dtype=np.uint64) + np.uint64((z + z0) * plane)
        out[z] = (splitmix64(i ^ np.uint64(seed)) < thr).reshape(ny, nx)
    return out


def _u01(seed, k):
    """k-th uniform [0,1) of stream `seed` (scalar seed, vector k)."""
    z = splitmix64((np.uint64(seed) * np.uint64(0x100000001B3) + np.asarray(k, dtype=np.uint64)) & _M64)
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def disc_image(index, n=128, base_seed=20240, rmin=6.0, rmax=24.0, max_discs=400):
    """One 'Mendeley-shaped' image: overlapping solid discs on a fluid background, drawn until a
    per-image target solid fraction ~ U(0.3, 0.8) is reached.  Returns u8 [n, n], 1 = solid."""
    seed = base_seed + index
    target = 0.3 + 0.5 * float(_u01(seed, 0))
    yy, xx = np.mgrid[0:n, 0:n]
    solid = np.zeros((n, n), dtype=bool)
    u = _u01(seed, np.arange(1, 3 * max_discs + 1))
    for d in range(max_discs):
        cx, cy, rr = u[3 * d] * n, u[3 * d + 1] * n, rmin + (rmax - rmin) * u[3 * d + 2]
        solid |= (xx + 0.5 - cx) ** 2 + (yy + 0.5 - cy) ** 2 <= rr * rr
        if solid.mean() >= target:
            break
    return solid.astype(np.uint8)


def disc_images(count, n=128, base_seed=20240, start=0):
    return np.stack([disc_image(start + i, n, base_seed) for i in range(count)])


def to_gray(solid):
    """Solid mask -> the 8-bit grey image the reference reads (white = solid, black = fluid)."""
    return (np.asarray(solid) != 0).astype(np.uint8) * 255
