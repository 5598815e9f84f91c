"""Kernel breakdown of the painter alone on a FULL mesh (no pipeline, no spectrum buffers -- lean enough for ncu's
kernel replay at C4's size: mesh 34 GB + records 32 GB + catalogue 12 GB):

   python tools/paint_profile.py [--n-mesh 2048] [--n-part 1e9] [--order 4] [--reps 3] [--warmup 2] [--clustered]

Prints one JSON line: per-kernel CUDA-event times recorded by the library (jps_profile_*)."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
ap = argparse.ArgumentParser()
ap.add_argument("--n-mesh", type=int, default=2048); ap.add_argument("--n-part", type=float, default=1e9)
ap.add_argument("--order", type=int, default=4); ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--warmup", type=int, default=2); ap.add_argument("--tag", default="")
a = ap.parse_args()
from jax_powspec_b200 import _lib, mas
dev = torch.device("cuda", 0); n, box = a.n_mesh, 2000.0
npart = int(a.n_part)
g = torch.Generator(device=dev); g.manual_seed(7)
x = torch.rand(npart, generator=g, device=dev) * box
y = torch.rand(npart, generator=g, device=dev) * box
z = torch.rand(npart, generator=g, device=dev) * box
mesh = torch.zeros((n, n, n), dtype=torch.float32, device=dev)
def step():
    mas.paint(mesh, x, y, z, None, 0.0, 0.0, 0.0, box, n, True, order=a.order, compat="fixed", method="sorted", inplace=True)
for _ in range(a.warmup): step()
torch.cuda.synchronize()
_lib.profile_reset(); _lib.profile_enable(True)
torch.cuda.nvtx.range_push("jps_timed")
for _ in range(a.reps): step()
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
prof = _lib.profile_snapshot(); _lib.profile_enable(False)
total = float(mesh.sum(dtype=torch.float64))
print(json.dumps({"tag": a.tag, "env": {k: v for k, v in os.environ.items() if k.startswith("JPS_")}, "n_mesh": n, "n_part": npart,
                  "order": a.order, "kernels_ms": {k: round(ms / c, 4) for k, (c, ms) in prof.items()},
                  "total_ms": round(sum(ms for _, (c, ms) in prof.items()) / a.reps, 3),
                  "mass_rel_err": abs(total / ((a.warmup + a.reps) * npart) - 1.0)}))
