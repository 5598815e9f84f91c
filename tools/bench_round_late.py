"""Quick GPU timings of the two late additions of round 1 (run on the GPU box, ~20 s):
  - the config-3 all-triangles sweep (bispec_pairs: 276 pairs x 20 angles at 256^3), cold and warm
  - the mock generator (row f-3) at the reference's own size (tests/create_lognormal.py:13-19)
Prints one JSON object."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import jax_powspec_b200 as jps
from jax_powspec_b200 import mocks

F32 = np.float32
res = {}


def gpu_time(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


n, box = 256, 1000.0
kf_t = np.linspace(1e-4, 10, 4056)
pk_t = 2.0e4 * (kf_t / 0.02) / (1.0 + (kf_t / 0.02) ** 2) ** 1.7
ms_field, dk = gpu_time(lambda: mocks.gaussian_field(n, kf_t, pk_t, 0, 100, box))
g = torch.fft.irfftn(dk, s=(n, n, n)).contiguous()
ms_pop, pos = gpu_time(lambda: mocks.populate_field(g, n, box, 3.5e-3, 101, lognormal_bias=1.1))
ms_mock, pos = gpu_time(lambda: mocks.lognormal_mock(n, kf_t, pk_t, 1.1, 3.5e-3, 100, box))
res["mock_256"] = {"gaussian_field_ms": ms_field, "populate_ms": ms_pop, "lognormal_mock_ms": ms_mock,
                   "particles": int(pos.shape[0])}
print(json.dumps(res), flush=True)

rho = jps.cic_mas_vec(torch.zeros((n, n, n), device="cuda"), pos[:, 0], pos[:, 1], pos[:, 2], None, pos.shape[0],
                      0., 0., 0., box, n, True)
delta = rho / rho.mean() - 1.0
theta = np.linspace(0, np.pi, 20).astype(F32)
kF = 2 * np.pi / box
centres = np.arange(2 * kF, 0.3, 2 * kF).astype(F32)
k1s, k2s = jps.triangle_pairs(centres)
torch.cuda.synchronize(); t0 = time.perf_counter()
sweep = jps.bispec_pairs(delta, box, k1s, k2s, theta)
torch.cuda.synchronize(); t_cold = time.perf_counter() - t0
t0 = time.perf_counter()
sweep = jps.bispec_pairs(delta, box, k1s, k2s, theta)
torch.cuda.synchronize(); t_warm = time.perf_counter() - t0
res["C3_all_triangles"] = {"pairs": int(k1s.size), "angles": int(theta.size), "triangle_bins": int(k1s.size * theta.size),
                           "cold_s": t_cold, "warm_s": t_warm, "finite_B": int(np.isfinite(sweep[3].cpu().numpy()).sum())}
print(json.dumps(res), flush=True)
