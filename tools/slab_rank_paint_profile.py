"""Kernel breakdown of ONE rank's paint stage of the slab-sharded C4 job (no communication needed):
   python tools/slab_rank_paint_profile.py [--world 8] [--n-mesh 2048] [--n-part 1e9] [--order 4]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
ap = argparse.ArgumentParser()
ap.add_argument("--world", type=int, default=8); ap.add_argument("--n-mesh", type=int, default=2048)
ap.add_argument("--n-part", type=float, default=1e9); ap.add_argument("--order", type=int, default=4)
a = ap.parse_args()
import jax_powspec_b200 as jps
from jax_powspec_b200 import _lib
from jax_powspec_b200.slab import SlabPipeline
dev = torch.device("cuda", 0); n, box = a.n_mesh, 2000.0
nloc = int(a.n_part) // a.world
g = torch.Generator(device=dev); g.manual_seed(42)
w_slab = box / a.world
x = torch.rand(nloc, generator=g, device=dev) * w_slab * 0.9999
y = torch.rand(nloc, generator=g, device=dev) * box; z = torch.rand(nloc, generator=g, device=dev) * box
ke = np.arange(2 * np.pi / box, np.pi * n / box, 2 * np.pi / box).astype(np.float32)
pipe = SlabPipeline(n, box, ke, order=a.order, compat="fixed", rank=0, world=a.world, transport="nccl")
for _ in range(3): pipe.stage_paint(x, y, z)
torch.cuda.synchronize()
_lib.profile_reset(); _lib.profile_enable(True)
for _ in range(5): pipe.stage_paint(x, y, z)
torch.cuda.synchronize()
prof = _lib.profile_snapshot(); _lib.profile_enable(False)
print(json.dumps({"n_part_rank": nloc, "tiles": None, "kernels_ms": {k: round(ms / c, 4) for k, (c, ms) in prof.items()},
                  "total_ms": round(sum(ms for _, (c, ms) in prof.items()) / 5, 3)}))
