"""Small driver for compute-sanitizer (tools/sanitize.sh): the bucketed tile painter (fixed-point deposit,
TMA reduce flush + per-thread red flush on boundary tiles), the big-mesh tile count (global reds) and the
slab peer-store kernels (plain, transposing, TMA bulk-store) with virtual ranks on one device -- each
checked against the plain atomic painter / the tensor-copy exchange so that a sanitizer-clean run is also
a correct one."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import jax_powspec_b200 as jps
from jax_powspec_b200.slab import SlabPipeline, run_virtual_ranks

rng = np.random.default_rng(3)
box = 1000.0
for n, npart in ((64, 60_000), (80, 40_000), (50, 20_000)):        # 16 | n, 16 !| n, n % 4 != 0
    p = (rng.random((npart, 3)) ** 1.3 * box).astype(np.float32)
    p[p >= np.float32(box)] = 0.0
    w = (rng.random(npart).astype(np.float32) - np.float32(0.3))
    for order, compat in ((2, "reference"), (2, "fixed"), (3, "fixed"), (4, "fixed")):
        for wt in (None, w):
            a = jps.paint(np.zeros((n, n, n), np.float32), p[:, 0], p[:, 1], p[:, 2], wt, 0., 0., 0., box, n, True,
                          order=order, compat=compat, method="atomic")
            b = jps.paint(np.zeros((n, n, n), np.float32), p[:, 0], p[:, 1], p[:, 2], wt, 0., 0., 0., box, n, True,
                          order=order, compat=compat, method="sorted")
            assert np.abs(a - b).max() <= 4e-6 * max(np.abs(a).max(), 1.0), (n, order, compat)
print("tile painter ok")
# big-mesh path: more tiles than the shared-memory histogram holds (N > 576)
n, npart = 592, 200_000
p = (rng.random((npart, 3)) * box).astype(np.float32)
p[p >= np.float32(box)] = 0.0
x, y, z = (torch.from_numpy(np.ascontiguousarray(p[:, i])).cuda() for i in range(3))
zero = torch.zeros((n, n, n), device="cuda")
a = jps.paint(zero, x, y, z, None, 0., 0., 0., box, n, True, order=4, compat="fixed", method="atomic")
b = jps.paint(zero, x, y, z, None, 0., 0., 0., box, n, True, order=4, compat="fixed", method="sorted")
assert float((a - b).abs().max()) <= 4e-6
print("big-mesh bucketing ok")
del a, b, zero
# slab peer-store kernels, virtual ranks on one device
n, npart = 64, 50_000
p = (rng.random((npart, 3)) * box).astype(np.float32)
p[p >= np.float32(box)] = 0.0
ke = np.arange(2 * np.pi / box, np.pi * n / box, 2 * np.pi / box).astype(np.float32)
for world in (2, 4):
    owner = (np.floor(p[:, 0] * np.float32(n / box)).astype(np.int64) % n) // (n // world)
    cats = [tuple(torch.from_numpy(np.ascontiguousarray(p[owner == r][:, i])).cuda() for i in range(3)) + (None,) for r in range(world)]
    ref = [SlabPipeline(n, box, ke, order=3, rank=r, world=world) for r in range(world)]
    k0, pk0, nm0 = run_virtual_ranks(ref, cats)[0]
    for layout, fft in (("xslow", "auto"), ("xfast", "cufft2d"), ("xfast", "pencil")):
        pipes = [SlabPipeline(n, box, ke, order=3, rank=r, world=world, fft=fft) for r in range(world)]
        for q in pipes:
            q._force_chunks = True
        k1, pk1, nm1 = run_virtual_ranks(pipes, cats, p2p=layout)[0]
        assert torch.equal(nm0, nm1) and float(((pk1 - pk0).abs() / pk0[:, :1].abs()).max()) < 2e-6, (world, layout, fft)
print("slab peer-store ok")
