"""CUDA P(k) path (cuFFT + fused binning kernel, through the C ABI) against the golden vectors
produced by the reference run under the JAX shim, and against the f64 oracle.

Tolerances: mode counts bit-exact; P0/P2/P4 within 1e-5 of |P0| per bin."""
import os

import numpy as np
import pytest
import torch

from oracle import correlations as oc
from oracle import mas as om
from tests.util import F32, clustered_particles, rel_to_monopole

pytestmark = pytest.mark.gpu

TOL = 1e-5


@pytest.fixture(scope="module")
def jps():
    import jax_powspec_b200
    return jax_powspec_b200


def _check_pk(got_pk, got_nm, want_pk, want_counts, tol=TOL):
    """Counts bit-exact; |dP_l|/|P0| <= tol per bin.  The forward FFT runs in float32 (as the
    reference's does): its error is ~1e-7 of the rms mode amplitude, i.e. an ABSOLUTE error set by
    the strongest modes, so a bin holding a handful of modes far below the peak power (e.g. the
    single Nyquist-corner mode) is only required to meet tol relative to that noise floor."""
    np.testing.assert_array_equal(got_nm.astype(np.int64), np.asarray(want_counts).astype(np.int64))
    want_counts = np.asarray(want_counts)
    ok = want_counts > 0
    assert np.all(np.isnan(got_pk[~ok])), "empty bins must be NaN like the reference (Q11)"
    want = np.asarray(want_pk)[ok].astype(np.float64)
    err = rel_to_monopole(got_pk[ok].astype(np.float64), want)
    p0 = np.abs(want[:, 0])
    floor = 1e-6 * np.sqrt(p0.max() / p0) / np.sqrt(want_counts[ok])      # float32 FFT noise, amplitude -> power
    allowed = tol + 9.0 * floor[:, None]                                   # 9 = largest (2l+1) L_l prefactor
    assert np.all(err <= allowed), f"max |dP|/|P0| = {err.max():.3e} (worst excess {np.max(err / allowed):.2f}x)"
    strong = (p0 >= 1e-2 * p0.max()) & (want_counts[ok] >= 8)
    if strong.any():
        assert err[strong].max() <= tol, f"max |dP|/|P0| on well-sampled bins = {err[strong].max():.3e}"


@pytest.mark.parametrize("tag", ["a", "b", "c"])
@pytest.mark.parametrize("edges", ["kf", "fine", "wide"])
def test_golden_powspec_vec(jps, golden_dir, tag, edges):
    g = np.load(os.path.join(golden_dir, f"ref_corr_{tag}.npz"))
    k3d, pk, nm = jps.powspec_vec(g["delta"], float(g["box"]), g[f"pk_{edges}_edges"])
    assert pk.dtype == F32 and nm.dtype == F32 and pk.shape == (len(nm), 3)
    np.testing.assert_array_equal(k3d, g[f"pk_{edges}_k3D"])              # bin centres, float32 exact (Q10)
    # the golden vectors are float32-serial sums (what XLA CPU does); 2e-5 absorbs their own rounding
    _check_pk(pk, nm, g[f"pk_{edges}_Pk3D"], g[f"pk_{edges}_Nmodes3D"], tol=2e-5)
    # and the real bar, against exact arithmetic
    _, pk64, counts = oc.powspec(g["delta"], float(g["box"]), g[f"pk_{edges}_edges"], precision="f64")
    _check_pk(pk, nm, pk64, counts)


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_golden_fundamental(jps, golden_dir, tag):
    g = np.load(os.path.join(golden_dir, f"ref_corr_{tag}.npz"))
    k3d, pk, nm = jps.powspec_vec_fundamental(g["delta"], float(g["box"]))
    np.testing.assert_array_equal(nm, g["pkf_Nmodes3D"])
    ok = nm > 0
    np.testing.assert_allclose(k3d[ok], g["pkf_k3D"][ok], rtol=1e-6)       # the .set quirk (Q18)
    _, pk64, counts = oc.powspec_fundamental(g["delta"], float(g["box"]), precision="f64")
    _check_pk(pk, nm, pk64, counts)
    k3d_fixed, _, _ = jps.powspec_vec_fundamental(g["delta"], float(g["box"]), compat="fixed")
    kf, _, _ = oc.powspec_fundamental(g["delta"], float(g["box"]), precision="f64", compat="fixed")
    np.testing.assert_allclose(k3d_fixed[ok], kf[ok], rtol=1e-6)


@pytest.fixture(scope="module")
def field128():
    n, box, npart = 128, 1000.0, 2_000_000
    p = clustered_particles(77, npart, box)
    rho = om.paint(np.zeros((n, n, n)), p[:, 0], p[:, 1], p[:, 2], None, 0., 0., 0., box, n, True,
                   order=2, compat="reference", precision="f64")
    delta = (rho / rho.mean() - 1.0).astype(F32)
    return n, box, p, rho.astype(F32), delta


@pytest.mark.parametrize("mas_order", [2, 3, 4])
def test_oracle_n128_all_windows(jps, field128, mas_order):
    n, box, p, rho, delta = field128
    kF = 2 * np.pi / box
    ke = np.arange(kF, np.pi * n / box, kF).astype(F32)
    k3d, pk, nm = jps.powspec_vec(delta, box, ke, mas_order=mas_order)
    _, pk64, counts = oc.powspec(delta, box, ke, mas_order=mas_order, precision="f64")
    _check_pk(pk, nm, pk64, counts)
    # faithful_f32 distance, for the record (reference-like serial float32 sums)
    _, pk32, _ = oc.powspec(delta, box, ke, mas_order=mas_order, precision="f32")
    print("faithful_f32 vs f64:", rel_to_monopole(pk32.astype(np.float64), pk64).max())


def test_many_bins_global_accumulator_path(jps, field128):
    n, box, p, rho, delta = field128
    kF = 2 * np.pi / box
    ke = np.arange(1e-4, 1.2 * np.sqrt(3) * np.pi * n / box, 0.11 * kF).astype(F32)   # > 576 reachable bins
    k3d, pk, nm = jps.powspec_vec(delta, box, ke)
    assert (nm > 0).sum() > 576
    _, pk64, counts = oc.powspec(delta, box, ke, precision="f64")
    _check_pk(pk, nm, pk64, counts)


def test_huge_bin_count_global_reds_path(jps, field128):
    """> 8192 reachable bins: neither shared-memory accumulator layout fits, float64 global reds are used."""
    n, box = 192, 1000.0
    delta = np.random.default_rng(8).standard_normal((n, n, n)).astype(F32)
    kF = 2 * np.pi / box
    ke = np.arange(1e-4, 1.01 * np.sqrt(3) * np.pi * n / box, 0.004 * kF).astype(F32)
    k3d, pk, nm = jps.powspec_vec(delta, box, ke)
    assert (nm > 0).sum() > 8192
    _, pk64, counts = oc.powspec(delta, box, ke, precision="f64")
    _check_pk(pk, nm, pk64, counts)


def test_normalise_folds_density_contrast(jps, field128):
    n, box, p, rho, delta = field128
    ke = np.arange(0.003, np.pi * n / box, 0.0025).astype(F32)          # tests/voids.py:55
    a = jps.powspec_vec(delta, box, ke)
    b = jps.powspec_vec(rho, box, ke, normalise=True)
    np.testing.assert_array_equal(a[2], b[2])
    ok = a[2] > 0
    assert rel_to_monopole(b[1][ok].astype(np.float64), a[1][ok].astype(np.float64)).max() < TOL
    c = jps.powspec_vec(rho, box, ke, normalise=True, shot_noise=123.0)
    np.testing.assert_allclose(c[1][ok][:, 0], b[1][ok][:, 0] - 123.0, rtol=1e-5, atol=1e-3)
    np.testing.assert_array_equal(c[1][ok][:, 1:], b[1][ok][:, 1:])


def test_device_tensors_and_raw_outputs(jps, field128):
    n, box, p, rho, delta = field128
    ke = np.arange(0.01, 0.3, 0.01).astype(F32)
    d = torch.from_numpy(delta).cuda()
    k3d, pk, nm, (sums, counts) = jps.powspec_vec(d, box, ke, return_raw=True)
    assert pk.is_cuda and sums.dtype == torch.float64 and counts.dtype == torch.int64
    _, pk64, c64 = oc.powspec(delta, box, ke, precision="f64")
    np.testing.assert_array_equal(counts.cpu().numpy(), c64)
    _check_pk(pk.cpu().numpy(), nm.cpu().numpy(), pk64, c64)


def test_plane_wave_known_answer(jps):
    # delta = A cos(2 pi m.x / N): all power in |k|^2 = m.m, P0*Nmodes/V = 2 * (A N^3 / 2)^2 W^2 for the
    # two stored half-space modes... we check it lands in exactly one bin with the analytic amplitude.
    n, box, A = 32, 100.0, 0.01
    m = (3, 2, 1)
    g = np.arange(n)
    phase = 2 * np.pi * (m[0] * g[:, None, None] + m[1] * g[None, :, None] + m[2] * g[None, None, :]) / n
    delta = (A * np.cos(phase)).astype(F32)
    k3d, pk, nm, (sums, counts) = jps.powspec_vec(delta, box, np.arange(0.5, 20, 1.0).astype(F32) * (2 * np.pi / box),
                                                  return_raw=True)
    kbin = int(np.floor(np.sqrt(14.0) - 0.5))
    w = np.prod([(np.pi * mi / n / np.sin(np.pi * mi / n)) ** 2 for mi in m])
    expect = (A * n ** 3 / 2) ** 2 * w ** 2          # one stored mode (kz = 1 > 0: its conjugate is not stored)
    assert abs(sums[kbin, 0] - expect) / expect < 1e-5
    others = np.delete(sums[:, 0], kbin)
    assert np.all(np.abs(others) < 1e-6 * expect)


def test_mode_counts_brute_force_n256(jps):
    n, box = 256, 2500.0
    ke = np.arange(1e-4, 5, 0.2e-2).astype(F32)                          # tests/correlations.py:76 (C1)
    delta = torch.zeros((n, n, n), device="cuda")
    _, _, nm = jps.powspec_vec(delta, box, ke)
    nm = nm.cpu().numpy()
    ki = oc.k_index(n).astype(np.int64)
    k2 = ki[:, None, None] ** 2 + ki[None, :, None] ** 2 + ki[None, None, : n // 2 + 1] ** 2
    bins = oc.bin_index_from_k(np.sqrt(k2.astype(F32)).ravel(), oc.grid_edges(ke, box))
    want = np.bincount(bins[bins >= 0], minlength=len(ke) - 1)
    np.testing.assert_array_equal(nm.astype(np.int64), want)
    assert (want > 0).sum() == 278 and want.sum() == 8454143              # SURVEY.md a-5 probe


def test_linearity_and_fused_pipeline(jps, field128):
    n, box, p, rho, delta = field128
    kF = 2 * np.pi / box
    ke = np.arange(kF, np.pi * n / box, kF).astype(F32)
    _, pk1, nm1 = jps.powspec_vec(delta, box, ke)
    _, pk2, nm2 = jps.powspec_vec(2.0 * delta, box, ke)
    np.testing.assert_allclose(pk2, 4.0 * pk1, rtol=2e-6)
    # fused paint -> P(k) equals the two-step path through the oracle
    k3d, pkf, nmf = jps.paint_powspec(p[:, 0], p[:, 1], p[:, 2], None, 0., 0., 0., box, n, ke, order=2,
                                      compat="reference", method="atomic")
    _, pk64, counts = oc.powspec(delta, box, ke, precision="f64")
    _check_pk(pkf, nmf, pk64, counts, tol=2e-5)


def test_host_pipeline_chunked_overlap_equals_one_shot(jps, field128):
    """The public host-buffer call streams the catalogue in chunks on a copy stream; the result must
    equal the device-resident one-shot pipeline (counts exact, P(k) to float32 summation order)."""
    n, box, p, rho, delta = field128
    kF = 2 * np.pi / box
    ke = np.arange(kF, np.pi * n / box, kF).astype(F32)
    pipe = jps.PaintPowspec(n, box, ke, order=3, compat="fixed", method="sorted", n_part_max=len(p))
    host = jps.HostPipeline(pipe, len(p), n_chunks=5)
    xh, yh, zh = (torch.from_numpy(np.ascontiguousarray(p[:, i])).pin_memory() for i in range(3))
    for _ in range(2):                                            # twice: buffer reuse across steps
        k3d, pk, nm = (a.copy() for a in host(xh, yh, zh))
    xd, yd, zd = (t.cuda() for t in (xh, yh, zh))
    k1, pk1, nm1 = (t.cpu().numpy() for t in pipe(xd, yd, zd))
    np.testing.assert_array_equal(nm, nm1)
    assert rel_to_monopole(pk.astype(np.float64), pk1.astype(np.float64)).max() < 2e-6
    rho64 = om.paint(np.zeros((n, n, n)), p[:, 0], p[:, 1], p[:, 2], None, 0., 0., 0., box, n, True,
                     order=3, compat="fixed", precision="f64")
    d64 = (rho64 / rho64.mean() - 1.0).astype(F32)
    _, pk64, counts = oc.powspec(d64, box, ke, mas_order=3, precision="f64")
    _check_pk(pk, nm, pk64, counts)


@pytest.mark.parametrize("n", [64, 96, 50, 130])
@pytest.mark.parametrize("order", [2, 4])
def test_pencil_fft_plan_equals_3d_plan_and_oracle(jps, n, order):
    """JPS_PLAN_FFT_PENCIL (three contiguous 1-D cuFFT passes + two transposing kernels, spectrum left as
    [kz][ky][kx], x-fastest binning kernel) against the monolithic 3-D plan and the f64 oracle."""
    box, npart = 1000.0, 150_000
    p = clustered_particles(500 + n, npart, box)
    kF = 2 * np.pi / box
    ke = np.arange(kF, np.pi * n / box, kF).astype(F32)
    x, y, z = (torch.from_numpy(np.ascontiguousarray(p[:, i])).cuda() for i in range(3))
    a = jps.PaintPowspec(n, box, ke, order=order, compat="fixed", fft="3d")
    b = jps.PaintPowspec(n, box, ke, order=order, compat="fixed", fft="pencil")
    assert a.fft == "3d" and b.fft == "pencil"
    ka, pka, nma = (t.cpu().numpy() for t in a(x, y, z))
    kb, pkb, nmb = (t.cpu().numpy() for t in b(x, y, z))
    np.testing.assert_array_equal(nma, nmb)
    np.testing.assert_array_equal(ka, kb)
    assert rel_to_monopole(pkb.astype(np.float64), pka.astype(np.float64)).max() < 5e-6
    rho = om.paint(np.zeros((n, n, n)), p[:, 0], p[:, 1], p[:, 2], None, 0., 0., 0., box, n, True,
                   order=order, compat="fixed", precision="f64")
    delta = (rho / rho.mean() - 1.0).astype(F32)
    _, pk64, counts = oc.powspec(delta, box, ke, mas_order=order, precision="f64")
    _check_pk(pkb, nmb, pk64, counts, tol=2e-5)
    # shot noise and a second call on the same plan (buffers reused)
    kb2, pkb2, _ = (t.cpu().numpy() for t in b(x, y, z))
    assert rel_to_monopole(pkb2.astype(np.float64), pkb.astype(np.float64)).max() < 2e-6     # atomic order of the painter only
