"""Run in a subprocess (JPS_COUNT / JPS_FINE / JPS_FINE_CHUNK are read once per process): the big-mesh bucketing path
(more tiles than the shared-memory histogram holds, N > 576) against the plain atomic painter."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import jax_powspec_b200 as jps
from tests.util import clustered_particles

n, box = 592, 1000.0
npart = 80_000 if "--quick" in sys.argv else 400_000
p = clustered_particles(11, npart, box)
w = (0.5 + np.random.default_rng(2).random(npart)).astype(np.float32)
x, y, z = (torch.from_numpy(np.ascontiguousarray(p[:, i])).cuda() for i in range(3))
wd = torch.from_numpy(w).cuda()
zero = torch.zeros((n, n, n), device="cuda")
for order in (3, 4):
    for wt in (None, wd):
        a = jps.paint(zero, x, y, z, wt, 0., 0., 0., box, n, True, order=order, compat="fixed", method="atomic")
        b = jps.paint(zero, x, y, z, wt, 0., 0., 0., box, n, True, order=order, compat="fixed", method="sorted")
        err = float((a - b).abs().max()) / max(float(a.abs().max()), 1.0)
        assert err <= 4e-6, (order, wt is not None, err)
        tot = float(b.sum(dtype=torch.float64)); want = float(npart if wt is None else wd.sum(dtype=torch.float64))
        assert abs(tot / want - 1.0) < 1e-6, (order, tot, want)
print("big-mesh ok")
