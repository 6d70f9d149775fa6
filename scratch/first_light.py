import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import keff_cfd_b200 as K
from oracle import refbind as R

G = 'tests/golden/'
img = np.load(G + 'img_00009.npy')
solid = K.structure_from_image(img)
rng = np.random.default_rng(0)

# operator parity 2D
A, b, _ = R.orc_assemble2d(img)
x = rng.standard_normal(solid.shape)
for flags in (0, K.F_NO_TMA):
    p = K.default_params(flags=flags)
    y, bb = K.apply_operator(solid, x, p, want_b=True)
    yo = R.orc_apply(A, (1,) + solid.shape, x.ravel()).reshape(solid.shape)
    print('2D apply flags', flags, 'max rel err', np.abs(y - yo).max() / np.abs(yo).max(), 'b err', np.abs(bb.ravel() - b).max())

# operator parity 3D, odd sizes too
for shape in ((12, 12, 14), (9, 20, 70), (5, 7, 9), (40, 24, 130)):
    m = (rng.random(shape) < 0.4).astype(np.uint8)
    A, b, _ = R.orc_assemble3d(m, TL=0.3, TR=2.0)
    x = rng.standard_normal(shape)
    yo = R.orc_apply(A, shape, x.ravel()).reshape(shape)
    for flags in (0, K.F_NO_TMA):
        p = K.default_params(flags=flags, tl=0.3, tr=2.0)
        y, bb = K.apply_operator(m, x, p, want_b=True)
        print('3D apply', shape, 'flags', flags, 'err', np.abs(y - yo).max() / np.abs(yo).max(), 'b err', np.abs(bb.ravel() - b).max())

# 2D solve
for flags in (0, K.F_NO_TMA):
    t = time.time()
    r = K.solve2d(solid, K.default_params(flags=flags), want_q=True)
    print('2D solve flags', flags, {k: r[k] for k in ('keff', 'q_left', 'q_right', 'rel_residual', 'energy_residual', 'iters', 'status', 'gpu_ms')}, 'wall', time.time() - t)
d = R.orc_discrete(img)
print('   oracle discrete', d['keff'], 'T maxdiff', np.abs(r['T'] - d['T']).max())

# 3D Schwarz
for name in ('schwarz163', 'schwarz091'):
    m = np.unpackbits(np.load(G + name + '_bits.npy')).reshape(128, 128, 128)
    t = time.time()
    r = K.solve3d(m)
    print(name, {k: r[k] for k in ('keff', 'q_left', 'q_right', 'rel_residual', 'iters', 'status', 'gpu_ms')}, 'wall', time.time() - t)

# batch
imgs = np.stack([solid, K.structure_from_image(np.load(G + 'img_00000.npy'))] * 4)
t = time.time()
rs = K.solve2d_batch(imgs)
print('batch', [(round(r['keff'], 10), r['iters'], r['status'], r['rel_residual']) for r in rs[:2]], 'gpu_ms', rs[0]['gpu_ms'], 'wall', time.time() - t)

# bigger 3D + stencil timing
import torch
for n in (256, 512):
    g = torch.Generator(device='cuda'); g.manual_seed(1)
    md = (torch.rand((n, n, n), device='cuda', generator=g) < 0.5).to(torch.uint8)
    for flags in (0, K.F_NO_TMA):
        ms = K.api.bench_apply(md, K.default_params(flags=flags), reps=20)
        print(f'apply {n}^3 flags {flags}: {ms:.4f} ms  {n**3/ms/1e6:.1f} Gcell/s  {16*n**3/ms/1e6:.0f} GB/s')
    t = time.time()
    r = K.solve3d(md, want_T=False)
    print(f'solve {n}^3', {k: r[k] for k in ('keff', 'q_left', 'q_right', 'rel_residual', 'iters', 'status', 'gpu_ms')}, 'wall', time.time() - t)
print('launches', K.lib().keffb200_launch_count())
