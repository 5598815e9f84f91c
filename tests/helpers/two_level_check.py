"""Run in a subprocess with JPS_BUCKET=two (and, optionally, JPS_FINE / JPS_MAX_GROUPS): the two-level partition
against the f64 oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import jax_powspec_b200 as jps
from oracle import mas as om
from tests.util import clustered_particles

assert os.environ.get("JPS_BUCKET") == "two"
F32 = np.float32
quick = "--quick" in sys.argv                      # compute-sanitizer runs (tools/sanitize.sh)
cases = ((50, 600.0, 40_000), (256, 2500.0, 250_000)) if quick else ((64, 1000.0, 200_000), (50, 600.0, 60_000), (256, 2500.0, 600_000))
for n, box, npart in cases:
    p = clustered_particles(3, npart, box)
    w = (0.5 + np.random.default_rng(1).random(npart)).astype(F32)
    for order, compat in (((2, "reference"), (4, "fixed")) if quick else ((2, "reference"), (2, "fixed"), (3, "fixed"), (4, "fixed"))):
        want = om.paint(np.zeros((n, n, n)), p[:, 0], p[:, 1], p[:, 2], w, 0., 0., 0., box, n, True,
                        order=order, compat=compat, precision="f64")
        got = jps.paint(np.zeros((n, n, n), F32), p[:, 0], p[:, 1], p[:, 2], w, 0., 0., 0., box, n, True,
                        order=order, compat=compat, method="sorted")
        err = np.abs(got - want) / np.maximum(np.abs(want), want.mean())
        assert err.max() < 4e-6, (n, order, compat, err.max())
print("two-level ok")
