"""How the bucketed painter degrades on strongly clustered catalogues (one CTA deposits one tile, heavy tiles are not split):
   python tools/clustered_paint.py      C2's size (1e8 particles, TSC, 512^3), three catalogues
Prints one JSON line per catalogue: particles in the densest tile, per-kernel times, mass conservation."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from jax_powspec_b200 import _lib, mas
dev = torch.device("cuda", 0); n, box, npart, order = 512, 2000.0, 100_000_000, 3
g = torch.Generator(device=dev); g.manual_seed(11)

def catalogue(frac_blobs, n_blobs, sigma_frac):
    nb = int(npart * frac_blobs)
    p = torch.rand((npart, 3), generator=g, device=dev) * box
    if nb:
        centres = torch.rand((n_blobs, 3), generator=g, device=dev) * box
        which = torch.randint(0, n_blobs, (nb,), generator=g, device=dev)
        p[:nb] = centres[which] + torch.randn((nb, 3), generator=g, device=dev) * (sigma_frac * box)
        p.remainder_(box)
        p[p >= box] = 0.0
    return p[torch.randperm(npart, generator=g, device=dev)].contiguous()

mesh = torch.zeros((n, n, n), dtype=torch.float32, device=dev)
for name, fb, nbl, sig in (("uniform", 0.0, 1, 0.0), ("half in 40 blobs of sigma = 2 % of the box", 0.5, 40, 0.02),
                           ("half in 4 blobs of sigma = 1 % of the box", 0.5, 4, 0.01)):
    p = catalogue(fb, nbl, sig)
    x, y, z = p[:, 0], p[:, 1], p[:, 2]
    tile = ((x * (n / box)).long().clamp_(0, n - 1) // 16 * 32 + (y * (n / box)).long().clamp_(0, n - 1) // 16) * 32 + (z * (n / box)).long().clamp_(0, n - 1) // 16
    heaviest = int(torch.bincount(tile, minlength=32768).max())
    del tile
    def step():
        mesh.zero_()
        mas.paint(mesh, x, y, z, None, 0.0, 0.0, 0.0, box, n, True, order=order, compat="fixed", method="sorted", inplace=True)
    for _ in range(2): step()
    torch.cuda.synchronize()
    _lib.profile_reset(); _lib.profile_enable(True)
    for _ in range(3): step()
    torch.cuda.synchronize()
    prof = _lib.profile_snapshot(); _lib.profile_enable(False)
    print(json.dumps({"catalogue": name, "densest_tile": heaviest, "mean_per_tile": npart / 32768,
                      "kernels_ms": {k: round(ms / c, 4) for k, (c, ms) in prof.items()},
                      "mass_rel_err": abs(float(mesh.sum(dtype=torch.float64)) / npart - 1.0)}), flush=True)
    del p, x, y, z
