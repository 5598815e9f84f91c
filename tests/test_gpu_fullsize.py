"""BASELINE.json configs[1] at its FULL size (1e8 particles, TSC, 512^3, kF-wide bins) through
size-independent properties: the oracle cannot run this in seconds, invariants can."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

N, BOX, NPART = 512, 2000.0, 100_000_000


@pytest.fixture(scope="module")
def c2():
    import jax_powspec_b200 as jps
    from jax_powspec_b200.mocks import lognormal_catalog
    x, y, z = lognormal_catalog(NPART, BOX, n_grid=256, seed=5, device="cuda")
    kF = 2.0 * math.pi / BOX
    ke = np.arange(kF, math.pi * N / BOX, kF).astype(np.float32)
    return jps, x, y, z, ke


def test_mass_linearity_and_permutation(c2):
    jps, x, y, z, ke = c2
    zero = torch.zeros((N, N, N), device="cuda")
    whole = jps.tsc_mas_vec(zero, x, y, z, None, NPART, 0., 0., 0., BOX, N, True)
    # sum of TSC weights is 1 per particle
    assert abs(whole.sum(dtype=torch.float64).item() - NPART) < 1e-6 * NPART
    # deposit is linear in the catalogue: first half, then the second half accumulated on top
    h = NPART // 2
    half = jps.tsc_mas_vec(zero, x[:h], y[:h], z[:h], None, h, 0., 0., 0., BOX, N, True)
    both = jps.tsc_mas_vec(half, x[h:], y[h:], z[h:], None, NPART - h, 0., 0., 0., BOX, N, True)
    tol = 4e-6 * torch.clamp(whole.abs(), min=1.0)
    assert bool(((both - whole).abs() <= tol).all())
    # and does not depend on the order of the particles (fixed-point tile sums; float32 only in the flush)
    perm = torch.randperm(NPART, device="cuda")
    shuffled = jps.tsc_mas_vec(zero, x[perm], y[perm], z[perm], None, NPART, 0., 0., 0., BOX, N, True)
    del perm
    assert bool(((shuffled - whole).abs() <= tol).all())
    # weights: w = 2 everywhere doubles the mesh exactly (powers of two commute with every rounding)
    w2 = torch.full((NPART,), 2.0, device="cuda")
    doubled = jps.tsc_mas_vec(zero, x, y, z, w2, NPART, 0., 0., 0., BOX, N, True)
    assert bool(((doubled - 2.0 * whole).abs() <= 2.0 * tol).all())


def test_fused_pipeline_equals_the_separate_calls_and_counts_every_mode(c2):
    jps, x, y, z, ke = c2
    zero = torch.zeros((N, N, N), device="cuda")
    rho = jps.tsc_mas_vec(zero, x, y, z, None, NPART, 0., 0., 0., BOX, N, True)
    k_a, pk_a, nm_a = jps.powspec_vec(rho / rho.mean() - 1.0, BOX, ke, mas_order=3)
    k_b, pk_b, nm_b = jps.paint_powspec(x, y, z, None, 0., 0., 0., BOX, N, ke, order=3)
    assert torch.equal(nm_a, nm_b) and torch.equal(k_a, k_b)
    # same density to ~1e-7 (mean folded through the DC mode vs divided in float32) -> same multipoles
    scale = pk_a[:, 0].abs()
    assert bool(((pk_a - pk_b).abs() <= 2e-4 * scale[:, None] + 1e-3).all())
    # mode counts are geometry: every stored half-space mode with kF <= |k| < last edge, once (Q7)
    ki = torch.fft.fftfreq(N, d=1.0 / N, device="cuda").to(torch.int64)
    ki[N // 2] = N // 2
    kz = torch.arange(N // 2 + 1, device="cuda", dtype=torch.int64)
    k2 = (ki[:, None, None] ** 2 + ki[None, :, None] ** 2 + kz[None, None, :] ** 2)
    kf = torch.sqrt(k2.to(torch.float32))
    kF = np.float32(2.0 * np.pi) / np.float32(BOX)
    e = torch.as_tensor((ke / kF).astype(np.float32), device="cuda")
    bins = torch.bucketize(kf, e, right=True) - 1
    bins = torch.where(kf == e[-1], torch.full_like(bins, len(e) - 2), bins)
    ok = (bins >= 0) & (bins < len(e) - 1)
    want = torch.bincount(bins[ok], minlength=len(e) - 1).to(torch.float32)
    assert torch.equal(nm_a, want)
    # shot-noise sanity at the smallest scales: P0 ~ 1/nbar within a factor of a few at k_Nyquist/1
    nbar = NPART / BOX ** 3
    assert 0.2 / nbar < float(pk_a[-1, 0]) < 5.0 / nbar


# =====================================================================================================
# Oracle parity AT BASELINE's own sizes: C1 and C3 in full, C2 one realisation.
# The checker is the float64 restatement -- NumPy for the estimators, its C + OpenMP twin
# (oracle/cport.py paint_f64 / powspec_f64, pinned on the NumPy oracle in tests/test_cport_cpu.py) where
# NumPy would need minutes.  Numbers are also written to gpurun_out/fullsize_parity.json.
# =====================================================================================================
import json
import os

from oracle import cport
from oracle import correlations as oc
from tests.util import rel_to_monopole

_REPORT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "fullsize_parity.json")


def _report(key, payload):
    try:
        os.makedirs(os.path.dirname(_REPORT), exist_ok=True)
        data = {}
        if os.path.exists(_REPORT):
            with open(_REPORT) as f:
                data = json.load(f)
        data[key] = payload
        with open(_REPORT, "w") as f:
            json.dump(data, f, indent=1)
    except OSError:
        pass
    print(key, json.dumps(payload))


def _pk_errors(pk, nm, pk64, c64):
    """Counts must be bit-equal; returns per-bin |dP_l|/|P0| against the float64 oracle and the bins over 1e-5."""
    np.testing.assert_array_equal(nm.astype(np.int64), c64)
    ok = c64 > 0
    err = rel_to_monopole(pk[ok].astype(np.float64), pk64[ok])
    over = [{"bin": int(b), "modes": int(c64[ok][i]), "P0": float(pk64[ok][i, 0]), "err": [float(v) for v in err[i]]}
            for i, b in enumerate(np.where(ok)[0]) if err[i].max() > 1e-5]
    return ok, err, over


def test_c1_full_size_reference_flow_against_f64_oracle():
    """BASELINE.json configs[0] as /root/reference/tests/correlations.py:41-78 runs it: 5e6 particles,
    ``cic_mas_vec`` (reference compat, Q1-Q4) on 256^3, ``delta /= delta.mean(); delta -= 1``,
    ``powspec_vec`` with ``arange(1e-4, 5, 0.2e-2)`` edges (2499 bins, 278 non-empty)."""
    import jax_powspec_b200 as jps
    from jax_powspec_b200.mocks import lognormal_catalog
    n, box, npart = 256, 2500.0, 5_000_000
    x, y, z = lognormal_catalog(npart, box, n_grid=128, seed=1005638091 % 2 ** 31, device="cuda")
    w = torch.ones(npart, device="cuda")
    rho = jps.cic_mas_vec(torch.zeros((n, n, n), device="cuda"), x, y, z, w, npart, 0., 0., 0., box, n, True)
    xh, yh, zh = (t.cpu().numpy() for t in (x, y, z))
    rho64 = cport.paint_f64(np.zeros((n, n, n)), xh, yh, zh, None, 0., 0., 0., box, n, True, order=2, compat="reference")
    mesh_err = np.abs(rho.cpu().numpy().astype(np.float64) - rho64) / np.maximum(np.abs(rho64), rho64.mean())
    assert mesh_err.max() <= 4e-6, f"mesh differs from the f64 oracle by {mesh_err.max():.2e} of max(|cell|, mean)"
    delta = rho / rho.mean()
    delta -= 1.0
    ke = np.arange(1e-4, 5, 0.2e-2).astype(np.float32)
    k3d, pk, nm = (t.cpu().numpy() for t in jps.powspec_vec(delta, box, ke))
    k64, pk64, c64 = cport.powspec_f64(delta.cpu().numpy(), box, ke)
    np.testing.assert_array_equal(k3d, k64)
    ok, err, over = _pk_errors(pk, nm, pk64, c64)
    assert int(ok.sum()) == 278 and int(c64.sum()) == 8_454_143          # every stored mode but DC
    assert np.all(np.isnan(pk[~ok]))                                     # Q11
    # the same catalogue through the reference-like serial float32 sums, for the record: ITS distance to
    # exact arithmetic is what a 1e-5 bar against "the reference" can mean (measured 1.5e-5)
    _, pk32, _ = cport.powspec(delta.cpu().numpy(), box, ke)
    err32 = rel_to_monopole(pk32[ok].astype(np.float64), pk64[ok])
    _report("C1", {"mesh_max_err": float(mesh_err.max()), "max_rel_P0": [float(v) for v in err.max(axis=0)],
                   "bins_over_1e-5": over, "faithful_f32_port_max_rel_P0": [float(v) for v in err32.max(axis=0)],
                   "nonempty_bins": int(ok.sum()), "modes": int(c64.sum())})
    assert err.max() <= 1e-5, f"|dP|/P0 = {err.max():.2e} against the f64 oracle; bins over 1e-5: {over}"


def test_c3_full_size_bispectrum_call_and_sweep_rows_against_f64_oracle():
    """BASELINE.json configs[2]: the reference's own call (/root/reference/tests/bispec.py:53-56:
    k1 = 0.1, k2 = 0.2, 20 angles) on a 256^3 CIC mesh of a 3.5e6-particle lognormal mock, and 20 rows of the
    all-triangles sweep (shells centred at 2 kF j < 0.3 h/Mpc, 20 angles each) against the float64 oracle."""
    import jax_powspec_b200 as jps
    from jax_powspec_b200.mocks import lognormal_catalog
    n, box, npart = 256, 1000.0, 3_500_000
    x, y, z = lognormal_catalog(npart, box, n_grid=128, seed=5, device="cuda")
    rho = jps.cic_mas_vec(torch.zeros((n, n, n), device="cuda"), x, y, z, None, npart, 0., 0., 0., box, n, True)
    delta = rho / rho.mean() - 1.0
    dh = delta.cpu().numpy()
    theta = np.linspace(0, np.pi, 20).astype(np.float32)
    stats = {}

    def compare(tag, got, want):
        ka, pk, _, B, Q = (np.asarray(t.cpu() if isinstance(t, torch.Tensor) else t) for t in got)
        ka64, pk64, _, B64, Q64 = want
        np.testing.assert_allclose(ka, ka64, rtol=3e-7)
        out = {}
        for name, g, w_ in (("P", pk, pk64), ("B", B, B64), ("Q", Q, Q64)):
            m = np.isfinite(w_)
            out[name] = float(np.abs(g[m] - w_[m]).max() / np.abs(w_[m]).max())
        stats[tag] = out
        assert max(out.values()) <= 1e-5, f"{tag}: {out}"

    compare("reference_call", jps.bispec(delta, box, 0.1, 0.2, theta), oc.bispec(dh, box, 0.1, 0.2, theta, precision="f64"))
    kF = 2 * np.pi / box
    centres = np.arange(2 * kF, 0.3, 2 * kF).astype(np.float32)
    k1s, k2s = jps.triangle_pairs(centres)
    assert k1s.size == 276                                               # 23 shell centres
    rows = np.linspace(0, k1s.size - 1, 20).round().astype(int)           # 20 rows spread over the sweep
    # every 4th angle of the sweep's 20 (the float64 oracle needs two 256^3 inverse FFTs per shell: 20 rows x
    # 20 angles would be four minutes of host time; the reference's own call above has all 20)
    th5 = np.ascontiguousarray(theta[::4])
    k_all, pk, _, B, Q = (t.cpu().numpy() for t in jps.bispec_pairs(delta, box, k1s[rows], k2s[rows], th5))
    for i, r in enumerate(rows):
        compare(f"sweep_row_{r}", (k_all[i], pk[i], th5, B[i], Q[i]),
                oc.bispec(dh, box, k1s[r], k2s[r], th5, precision="f64"))
    _report("C3", {"rows": [int(r) for r in rows], "max_rel": {k: max(s[k] for s in stats.values()) for k in "PBQ"},
                   "reference_call": stats["reference_call"]})


def test_c2_full_size_against_f64_oracle(c2):
    """BASELINE.json configs[1], one realisation at full size: 1e8 particles, TSC on 512^3, kF-wide bins --
    mesh and multipoles against the float64 C + OpenMP oracle (about half a minute of host time)."""
    jps, x, y, z, ke = c2
    rho = jps.tsc_mas_vec(torch.zeros((N, N, N), device="cuda"), x, y, z, None, NPART, 0., 0., 0., BOX, N, True)
    xh, yh, zh = (t.cpu().numpy() for t in (x, y, z))
    rho64 = cport.paint_f64(np.zeros((N, N, N)), xh, yh, zh, None, 0., 0., 0., BOX, N, True, order=3, compat="fixed")
    del xh, yh, zh
    rho_h = rho.cpu().numpy()
    mesh_err = float((np.abs(rho_h - rho64) / np.maximum(np.abs(rho64), rho64.mean())).max())
    assert mesh_err <= 4e-6, f"mesh differs from the f64 oracle by {mesh_err:.2e} of max(|cell|, mean)"
    del rho_h
    # the fused pipeline (bench.py's C2 step) against paint_f64 -> rho/mean - 1 -> float64 FFT -> float64 sums
    k3d, pk, nm = (t.cpu().numpy() for t in jps.paint_powspec(x, y, z, None, 0., 0., 0., BOX, N, ke, order=3))
    delta64 = rho64 / rho64.mean() - 1.0
    del rho64
    k64, pk64, c64 = cport.powspec_f64(delta64, BOX, ke, mas_order=3)
    ok, err, over = _pk_errors(pk, nm, pk64, c64)
    assert ok.all()
    _report("C2", {"mesh_max_err": mesh_err, "max_rel_P0": [float(v) for v in err.max(axis=0)], "bins_over_1e-5": over})
    assert err.max() <= 1e-5, f"|dP|/P0 = {err.max():.2e} against the f64 oracle; bins over 1e-5: {over}"
