"""Debug: atomic vs sorted TSC painter vs f64 oracle on a 512^3 mesh (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import jax_powspec_b200 as jps
from jax_powspec_b200.mocks import lognormal_catalog
from oracle import mas as om
n, box, npart = 512, 2000.0, int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
x, y, z = lognormal_catalog(npart, box, n_grid=128, seed=9, device="cuda")
zero = torch.zeros((n, n, n), device="cuda")
for order in (3,):
    a = jps.paint(zero, x, y, z, None, 0., 0., 0., box, n, True, order=order, compat="fixed", method="atomic")
    b = jps.paint(zero, x, y, z, None, 0., 0., 0., box, n, True, order=order, compat="fixed", method="sorted")
    b2 = jps.paint(zero, x, y, z, None, 0., 0., 0., box, n, True, order=order, compat="fixed", method="sorted")
    c = om.paint(np.zeros((n, n, n)), x.cpu().numpy(), y.cpu().numpy(), z.cpu().numpy(), None, 0., 0., 0., box, n, True,
                 order=order, compat="fixed", precision="f64")
    a, b, b2 = a.cpu().numpy().astype(np.float64), b.cpu().numpy().astype(np.float64), b2.cpu().numpy().astype(np.float64)
    for name, m in (("atomic", a), ("sorted", b), ("sorted2", b2)):
        d = np.abs(m - c)
        rel = d / np.maximum(np.abs(c), 1.0)
        i = np.unravel_index(np.argmax(rel), rel.shape)
        print(order, name, "max abs", d.max(), "max rel-to-max(|c|,1)", rel.max(), "at", i, "val", m[i], "oracle", c[i], "sum", m.sum(), c.sum())
    d = np.abs(a - b); i = np.unravel_index(np.argmax(d / np.maximum(np.abs(a), 1.0)), d.shape)
    print("a vs b worst", i, a[i], b[i], c[i], "n cells > 2e-5:", int((d > 2e-5 * np.maximum(np.abs(a), 1)).sum()))
    xs, ys, zs = (t.cpu().numpy() for t in (x, y, z))
    inv = np.float32(1.0) / (np.float32(box) / np.float32(n))
    px, py, pz = xs * inv, ys * inv, zs * inv
    near = (np.abs(px - i[0]) < 2.5) & (np.abs(py - i[1]) < 2.5) & (np.abs(pz - i[2]) < 2.5)
    print("particles near worst cell:", near.sum())
    for q in np.nonzero(near)[0][:12]:
        print("   ", repr(px[q]), repr(py[q]), repr(pz[q]))
