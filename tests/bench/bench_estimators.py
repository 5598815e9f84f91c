"""Timings of the other BASELINE configs on one GPU (run on the GPU box):
  C1: CIC (reference compat) on 256^3, 5e6 particles, P(k) with tests/correlations.py:76 edges
  C3: FFT bispectrum on a 256^3 CIC mesh, (k1,k2)=(0.1,0.2), 20 angles (tests/bispec.py:53-54); xi(s) too
  C5: covariance batch throughput (realisations/s) with 1e7 particles on 512^3 CIC
Prints one JSON object; CPU columns time oracle/ (C + NumPy restatement) on the host cores.
Lives under tests/ because it uses oracle/ as its checker (only tests/, smoke() and bench.py may)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import jax_powspec_b200 as jps
from jax_powspec_b200.mocks import lognormal_catalog
from jax_powspec_b200 import dist as jd
from oracle import cport, correlations as oc

F32 = np.float32
dev = torch.device("cuda", 0)
res = {}

def gpu_time(fn, reps=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out

# ---------------- C1
n, box, npart = 256, 2500.0, 5_000_000
x, y, z = lognormal_catalog(npart, box, n_grid=128, seed=1005638091 % 2**31, device=dev)
w = torch.ones(npart, device=dev)
ke = np.arange(1e-4, 5, 0.2e-2).astype(F32)
zero = torch.zeros((n, n, n), device=dev)
def c1():
    rho = jps.cic_mas_vec(zero, x, y, z, w, npart, 0., 0., 0., box, n, True)
    return jps.powspec_vec(rho, box, ke, normalise=True)
ms, (k3d, pk, nm) = gpu_time(c1)
xh, yh, zh = (t.cpu().numpy() for t in (x, y, z))
t0 = time.perf_counter()
rho_c = cport.paint(np.zeros((n, n, n), F32), xh, yh, zh, None, 0., 0., 0., box, n, True, order=2, compat="reference")
t1 = time.perf_counter()
d_c = rho_c / rho_c.mean() - F32(1)
kc, pkc, nmc = cport.powspec(d_c, box, ke)
t2 = time.perf_counter()
ok = nmc > 0
res["C1"] = {"gpu_ms": ms, "gparticles_per_s": npart / ms / 1e6, "cpu_paint_s": t1 - t0, "cpu_fft_bin_s": t2 - t1,
             "speedup_vs_cpu_port": (t2 - t0) * 1e3 / ms, "nonempty_bins": int(ok.sum()),
             "counts_equal": bool(np.array_equal(nm.cpu().numpy(), nmc)),
             "max_rel_to_P0": float(np.max(np.abs(pk.cpu().numpy()[ok] - pkc[ok]) / np.abs(pkc[ok][:, :1])))}

# ---------------- C3
n, box = 256, 1000.0
npart = 3_500_000
x, y, z = lognormal_catalog(npart, box, n_grid=128, seed=5, device=dev)
rho = jps.cic_mas_vec(torch.zeros((n, n, n), device=dev), x, y, z, None, npart, 0., 0., 0., box, n, True)
delta = rho / rho.mean() - 1.0
theta = np.linspace(0, np.pi, 20).astype(F32)
ms_b, out = gpu_time(lambda: jps.bispec(delta, box, 0.1, 0.2, theta), reps=5, warm=2)
se = np.arange(0.0, 200.0, 5.0).astype(F32)
ms_x, _ = gpu_time(lambda: jps.xi_vec(delta, box, se, guard_mu=True), reps=5, warm=2)
kk = np.arange(5e-3, 1, 2 * np.pi / box).astype(F32)
ms_all, _ = gpu_time(lambda: jps.compute_all_correlations(delta, box, se, kk, 0.1, 0.2, theta), reps=5, warm=2)
dh = delta.cpu().numpy()
t0 = time.perf_counter()
ref = oc.bispec(dh, box, 0.1, 0.2, theta, precision="f32")
t_cpu = time.perf_counter() - t0
B, B_ref = out[3].cpu().numpy(), np.asarray(ref[3], dtype=np.float64)
res["C3"] = {"bispec_gpu_ms": ms_b, "xi_gpu_ms": ms_x, "compute_all_gpu_ms": ms_all, "bispec_cpu_numpy_s": t_cpu,
             "speedup_vs_cpu_port": t_cpu * 1e3 / ms_b, "c2r_ffts_per_call": 44,
             "max_rel_B": float(np.max(np.abs(B - B_ref)) / np.max(np.abs(B_ref)))}

# ---------------- C3, the whole sweep: all (k1 <= k2) pairs of shells centred at 2 kF j up to 0.3 h/Mpc
if "--no-sweep" not in sys.argv:
    kF = 2 * np.pi / box
    centres = np.arange(2 * kF, 0.3, 2 * kF).astype(F32)
    k1s, k2s = jps.triangle_pairs(centres)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    sweep = jps.bispec_pairs(delta, box, k1s, k2s, theta)          # cold: indicator sums of every shell / triple computed
    torch.cuda.synchronize(); t_cold = time.perf_counter() - t0
    t0 = time.perf_counter()
    sweep = jps.bispec_pairs(delta, box, k1s, k2s, theta)          # warm: indicator sums cached on the device
    torch.cuda.synchronize(); t_warm = time.perf_counter() - t0
    res["C3_all_triangles"] = {"pairs": int(k1s.size), "angles": int(theta.size), "triangle_bins": int(k1s.size * theta.size),
                               "cold_s": t_cold, "warm_s": t_warm,
                               "finite_B": int(np.isfinite(sweep[3].cpu().numpy()).sum())}

# ---------------- mock generator (row f-3): the reference's recipe at its own size (tests/create_lognormal.py:13-19)
from jax_powspec_b200 import mocks
kf_t = np.linspace(1e-4, 10, 4056)
pk_t = 2.0e4 * (kf_t / 0.02) / (1.0 + (kf_t / 0.02) ** 2) ** 1.7
ms_field, dk_m = gpu_time(lambda: mocks.gaussian_field(256, kf_t, pk_t, 0, 100, 1000.0), reps=5, warm=2)
g_m = torch.fft.irfftn(dk_m, s=(256, 256, 256)).contiguous()
ms_pop, pos_m = gpu_time(lambda: mocks.populate_field(g_m, 256, 1000.0, 3.5e-3, 101, lognormal_bias=1.1), reps=5, warm=2)
ms_mock, pos_m = gpu_time(lambda: mocks.lognormal_mock(256, kf_t, pk_t, 1.1, 3.5e-3, 100, 1000.0), reps=5, warm=2)
res["mock_256"] = {"gaussian_field_ms": ms_field, "populate_ms": ms_pop, "lognormal_mock_ms": ms_mock,
                   "particles": int(pos_m.shape[0])}

# ---------------- C5 (one GPU): realisations/s
n, box, npart = 512, 1000.0, 10_000_000
ke = np.arange(0.003, np.pi * n / box, 0.0025).astype(F32)
pipe = jps.PaintPowspec(n, box, ke, order=2, compat="fixed", n_part_max=npart)
cats = [lognormal_catalog(npart, box, n_grid=128, seed=s, device=dev) for s in range(4)]
def measure(seed):
    xx, yy, zz = cats[seed % 4]
    return pipe(xx, yy, zz)[1].clone()
measure(0); torch.cuda.synchronize()
t0 = time.perf_counter()
rows, mean, cov = jd.covariance_batch(list(range(32)), measure)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
res["C5"] = {"realisations": 32, "seconds": dt, "realisations_per_s": 32 / dt, "cov_shape": list(cov.shape),
             "note": "paint+FFT+multipoles only (4 pre-generated catalogues cycled); mock generation excluded"}
# the same with every realisation GENERATED on the device first (lognormal_mock at 256^3, ~1e7 particles)
dens = npart / box ** 3
def measure_gen(seed):
    p = mocks.lognormal_mock(256, kf_t, pk_t, 1.1, dens, seed, box)
    m = min(p.shape[0], npart)
    return pipe(p[:m, 0], p[:m, 1], p[:m, 2])[1].clone()
measure_gen(0); torch.cuda.synchronize()
t0 = time.perf_counter()
rows, mean, cov = jd.covariance_batch(list(range(16)), measure_gen)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
res["C5_with_mock_generation"] = {"realisations": 16, "seconds": dt, "realisations_per_s": 16 / dt}
print(json.dumps(res))
