"""Row f-3 (mock generator): the per-mode / per-cell arithmetic the device kernels run
(jax_powspec_b200/csrc/mockgen.cuh), compiled for the host.  Known answers for the counter-based
generator, the reference's P(k) interpolation and Hermitian pairing
(/root/reference/src/gauss_field.py:5-80), the Poisson sampler and the triangular in-cell offsets
(/root/reference/src/populate_field.py:4-29)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
from scipy import stats

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "helpers", "mockgen_host.cpp")
OUT = os.path.join(HERE, "helpers", "_build", "libmockgen_host.so")
HDR = os.path.join(HERE, "..", "jax_powspec_b200", "csrc", "mockgen.cuh")

_dp = C.POINTER(C.c_double)


def load_host():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    if not os.path.exists(OUT) or os.path.getmtime(OUT) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", SRC, "-o", OUT])
    lib = C.CDLL(OUT)
    lib.mock_philox.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.mock_uniform_open.argtypes = [C.c_uint32, C.c_uint32]
    lib.mock_uniform_open.restype = C.c_double
    lib.mock_uniform_f32.argtypes = [C.c_uint32]
    lib.mock_uniform_f32.restype = C.c_float
    lib.mock_interp_power.argtypes = [_dp, _dp, C.c_int, C.c_double]
    lib.mock_interp_power.restype = C.c_double
    lib.mock_gaussian_field.argtypes = [C.c_int, _dp, _dp, C.c_int, C.c_int, C.c_ulonglong, C.c_double, C.c_void_p]
    lib.mock_field_uniforms.argtypes = [C.c_int, C.c_ulonglong, C.c_void_p]
    lib.mock_poisson_many.argtypes = [C.c_double, C.c_ulonglong, C.c_ulonglong, C.c_longlong, C.c_void_p]
    lib.mock_density_sum.argtypes = [C.c_void_p, C.c_longlong, C.c_int, C.c_double, C.c_int, C.c_int]
    lib.mock_density_sum.restype = C.c_double
    lib.mock_populate_count.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_int, C.c_double,
                                        C.c_ulonglong, C.c_double, C.c_void_p]
    lib.mock_populate_count.restype = C.c_longlong
    lib.mock_populate_fill.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_ulonglong, C.c_void_p]
    lib.mock_tri_offsets.argtypes = [C.c_void_p, C.c_longlong, C.c_float, C.c_void_p]
    return lib


@pytest.fixture(scope="module")
def host():
    return load_host()


def _dptr(a):
    return a.ctypes.data_as(_dp)


def host_gaussian_field(lib, n, kf, pkf, rayleigh, seed, box):
    kf = np.ascontiguousarray(kf, dtype=np.float64)
    pkf = np.ascontiguousarray(pkf, dtype=np.float64)
    out = np.zeros((n, n, n // 2 + 1, 2), dtype=np.float32)
    lib.mock_gaussian_field(n, _dptr(kf), _dptr(pkf), kf.size, int(rayleigh), seed, float(box), out.ctypes.data)
    return out[..., 0] + 1j * out[..., 1]


# ------------------------------------------------------------------ generator
def test_philox_known_answers(host):
    """Random123 known-answer vectors for philox4x32-10."""
    kat = [
        ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
        ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
        ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
         (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
    ]
    for ctr, key, want in kat:
        c, k, o = np.array(ctr, np.uint32), np.array(key, np.uint32), np.zeros(4, np.uint32)
        host.mock_philox(c.ctypes.data, k.ctypes.data, o.ctypes.data)
        assert tuple(int(v) for v in o) == want


def test_uniform_ranges(host):
    assert 0.0 < host.mock_uniform_open(0, 0) < 1e-15
    assert 1.0 - 1e-15 < host.mock_uniform_open(0xffffffff, 0xffffffff) < 1.0
    assert host.mock_uniform_f32(0) == 0.0
    assert host.mock_uniform_f32(0xffffffff) == np.float32(1.0) - np.float32(2.0 ** -24)


# ------------------------------------------------------------------ Gaussian field
def reference_interp(kf, pkf, kmod):
    """gauss_field.py:38-45, verbatim semantics."""
    lmin, lmax = 0, len(kf) - 1
    while lmax - lmin > 1:
        l = (lmin + lmax) // 2
        if kf[l] < kmod:
            lmin = l
        else:
            lmax = l
    return (pkf[lmax] - pkf[lmin]) / (kf[lmax] - kf[lmin]) * (kmod - kf[lmin]) + pkf[lmin]


def test_interpolation_matches_reference_bisection(host):
    rng = np.random.default_rng(3)
    kf = np.sort(rng.random(57)) + 0.01
    pkf = rng.random(57) * 1e4
    for kmod in list(rng.random(200) * 1.2) + [kf[0], kf[-1], kf[10], 0.0, 5.0]:
        got = host.mock_interp_power(_dptr(kf), _dptr(pkf), kf.size, float(kmod))
        assert got == reference_interp(kf, pkf, float(kmod))
    inside = rng.uniform(kf[0], kf[-1], 100)
    got = np.array([host.mock_interp_power(_dptr(kf), _dptr(pkf), kf.size, float(k)) for k in inside])
    np.testing.assert_allclose(got, np.interp(inside, kf, pkf), rtol=1e-12)


@pytest.mark.parametrize("n", [8, 12, 9])
def test_field_hermitian_and_amplitudes(host, n):
    box = 500.0
    kf = np.linspace(1e-3, 2.0, 400)
    pkf = 2.0e4 * (kf / 0.02) / (1.0 + (kf / 0.02) ** 2) ** 1.7
    dk = host_gaussian_field(host, n, kf, pkf, 0, 77, box)
    mid = n // 2
    assert dk[0, 0, 0] == 0
    freq = np.array([i - n if i > mid else i for i in range(n)])
    kx, ky, kz = np.meshgrid(freq, freq, freq[: mid + 1], indexing="ij")
    kmod = np.sqrt(kx ** 2 + ky ** 2 + kz ** 2) * 2 * np.pi / box
    want = np.interp(kmod, kf, pkf) * (n * n / box) ** 3
    got = np.abs(dk) ** 2
    m = np.ones_like(got, dtype=bool)
    m[0, 0, 0] = False
    np.testing.assert_allclose(got[m], want[m], rtol=2e-6)          # fixed amplitude: |delta_k|^2 == P(k)
    # pairing on the self-conjugate planes, exactly the reference's index arithmetic (gauss_field.py:27,31,69)
    for iz in (0, mid):
        for ix in range(n):
            for iy in range(n):
                mx = n - freq[ix] if freq[ix] > 0 else -freq[ix]
                my = n - freq[iy] if freq[iy] > 0 else -freq[iy]
                if (mx, my) == (ix, iy):
                    if (ix, iy, iz) != (0, 0, 0):
                        assert dk[ix, iy, iz].imag == 0 and dk[ix, iy, iz].real > 0
                else:
                    assert dk[mx, my, iz] == np.conj(dk[ix, iy, iz])
    if n % 2 == 0:
        # Hermitian half-space array -> the inverse transform loses nothing: forward again gives it back
        real = np.fft.irfftn(dk, s=(n, n, n), axes=(0, 1, 2))
        np.testing.assert_allclose(np.fft.rfftn(real), dk, atol=1e-4 * np.abs(dk).max())


def test_field_reference_fill_order(host):
    """Replay the reference's sequential loop (gauss_field.py:25-76) with OUR per-mode draws standing in
    for its stream: the array it builds must equal ours, i.e. the pairing rule is restated correctly."""
    n, box = 8, 300.0
    kf = np.linspace(1e-3, 3.0, 50)
    pkf = 1.0e3 / (1.0 + kf) ** 2
    ours = host_gaussian_field(host, n, kf, pkf, 1, 5, box)
    mid = n // 2
    # the draw of mode (ix,iy,iz) when it is its own canonical member = value at the first-visited member
    built = np.zeros_like(ours)
    for ix in range(n):
        kx = ix - n if ix > mid else ix
        mx = n - kx if kx > 0 else -kx
        for iy in range(n):
            ky = iy - n if iy > mid else iy
            my = n - ky if ky > 0 else -ky
            for iz in range(mid + 1):
                kz = iz
                if built[ix, iy, iz] == 0:
                    # own draw = what ours holds at the canonical (first visited) member
                    val = ours[ix, iy, iz]
                    built[ix, iy, iz] = val
                    if kz == 0 or kz == mid:
                        if built[mx, my, iz] == 0:
                            built[mx, my, iz] = np.conj(val)
                        if (mx, my) == (ix, iy):
                            built[ix, iy, iz] = abs(val)
    built[0, 0, 0] = 0
    np.testing.assert_array_equal(built, ours)


def test_field_rayleigh_statistics(host):
    n, box = 24, 400.0
    kf = np.linspace(1e-3, 3.0, 300)
    pkf = np.full_like(kf, 50.0)
    dk = host_gaussian_field(host, n, kf, pkf, 1, 1234, box)
    ratio = (np.abs(dk) ** 2 / (50.0 * (n * n / box) ** 3))[:, :, 1: n // 2].ravel()   # interior planes: independent
    # |delta_k|^2 / P = -log(u): exponential with unit mean
    assert abs(ratio.mean() - 1.0) < 5.0 / np.sqrt(ratio.size)
    assert stats.kstest(ratio, "expon").pvalue > 1e-3
    phase = np.angle(dk[:, :, 1: n // 2]).ravel()
    assert stats.kstest((phase + np.pi) / (2 * np.pi), "uniform").pvalue > 1e-3
    # different seeds, different fields; same seed, same field
    assert not np.array_equal(dk, host_gaussian_field(host, n, kf, pkf, 1, 1235, box))
    np.testing.assert_array_equal(dk, host_gaussian_field(host, n, kf, pkf, 1, 1234, box))


# ------------------------------------------------------------------ Poisson + offsets
@pytest.mark.parametrize("lam", [0.02, 0.7, 3.5, 11.9, 12.0, 47.3, 1.0e3, 2.5e5])
def test_poisson_distribution(host, lam):
    count = 400_000
    out = np.zeros(count, dtype=np.uint32)
    host.mock_poisson_many(lam, 99, 1 << 33, count, out.ctypes.data)       # 64-bit cell counters
    k = out.astype(np.int64)
    assert abs(k.mean() - lam) < 5.0 * np.sqrt(lam / count)
    assert abs(k.var() - lam) < 6.0 * lam * np.sqrt(2.0 / count) + 6.0 * np.sqrt(lam / count)
    # chi-square against the exact pmf on bins with expectation >= 20
    lo, hi = int(stats.poisson.ppf(1e-4, lam)), int(stats.poisson.ppf(1 - 1e-4, lam)) + 1
    edges = np.arange(lo, hi + 1)
    if edges.size > 60:                                                   # coarsen wide distributions
        edges = np.unique(np.linspace(lo, hi, 60).astype(np.int64))
    cdf = stats.poisson.cdf(edges - 1, lam)
    exp = np.diff(np.concatenate([[0.0], cdf, [1.0]])) * count            # (-inf, lo), [e_i, e_i+1), [hi, inf)
    obs = np.histogram(k, bins=np.concatenate([[-1], edges, [np.iinfo(np.int64).max]]))[0]
    keep = exp >= 20
    chi2 = ((obs[keep] - exp[keep]) ** 2 / exp[keep]).sum()
    assert stats.chi2.sf(chi2, keep.sum() - 1) > 1e-4, (lam, chi2, keep.sum())


def test_poisson_edge_rates(host):
    out = np.zeros(1000, dtype=np.uint32)
    for lam in (0.0, -1.0, float("nan")):
        out[:] = 7
        host.mock_poisson_many(lam, 1, 0, out.size, out.ctypes.data)
        assert not out.any()


def test_triangular_offsets(host):
    rng = np.random.default_rng(8)
    bits = rng.integers(0, 2 ** 32, 500_000, dtype=np.uint64).astype(np.uint32)
    out = np.zeros(bits.size, dtype=np.float32)
    b = np.float32(3.90625)
    host.mock_tri_offsets(bits.ctypes.data, bits.size, b, out.ctypes.data)
    # populate_field.py:4-9 in float32
    u = (bits >> 8).astype(np.float32) * np.float32(2.0 ** -24)
    r = np.float32(2.0) * u - np.float32(1.0)
    want = np.sign(r) * (np.float32(1.0) - np.sqrt(np.abs(r))) * b
    np.testing.assert_array_equal(out, want.astype(np.float32))
    assert np.abs(out).max() <= b
    assert abs(out.mean()) < 5 * b / np.sqrt(6 * bits.size)
    assert abs(out.var() - b * b / 6.0) < 0.01 * b * b                   # triangular on (-b, b): variance b^2/6


def test_populate_host_end_to_end(host):
    n, box, density = 16, 200.0, 0.02
    rng = np.random.default_rng(0)
    rho = np.exp(rng.normal(size=(n, n, n))).astype(np.float32)
    s = host.mock_density_sum(rho.ctypes.data, rho.size, 0, 0.0, 148 * 8, 256)
    assert abs(s - rho.astype(np.float64).sum()) < 1e-9 * s
    counts = np.zeros(rho.size, dtype=np.uint32)
    total = host.mock_populate_count(rho.ctypes.data, n, box, density, 0, 0.0, 21, s, counts.ctypes.data)
    assert total == counts.sum()
    expect = density * box ** 3
    assert abs(total - expect) < 5 * np.sqrt(expect)
    lam = rho.astype(np.float64).ravel() * (box / n) ** 3 * density / rho.astype(np.float64).mean()
    # counts follow the cell rates: Pearson residuals have unit variance
    resid = (counts - lam) / np.sqrt(lam)
    assert abs(resid.var() - 1.0) < 0.1 and abs(resid.mean()) < 0.05
    pos = np.zeros((total, 3), dtype=np.float32)
    host.mock_populate_fill(counts.ctypes.data, n, np.float32(box), 21, pos.ctypes.data)
    assert (pos >= 0).all() and (pos < np.float32(box)).all()
    # every particle within one cell size of its cell centre (periodic), grouped by cell in C order
    cell = np.repeat(np.arange(rho.size), counts)
    centre = (np.stack(np.unravel_index(cell, (n, n, n)), axis=1) + 0.5) * (box / n)
    d = np.abs(pos - centre)
    d = np.minimum(d, box - d)
    assert d.max() <= box / n * (1 + 1e-6)
    # lognormal switch: exp(bias * g) sampled directly from the Gaussian field
    g = rng.normal(size=(n, n, n)).astype(np.float32) * 0.5
    s2 = host.mock_density_sum(g.ctypes.data, g.size, 1, 1.5, 148 * 8, 256)
    assert abs(s2 - np.exp(1.5 * g.astype(np.float64)).sum()) < 1e-9 * s2
    total2 = host.mock_populate_count(g.ctypes.data, n, box, density, 1, 1.5, 22, s2, counts.ctypes.data)
    assert abs(total2 - expect) < 5 * np.sqrt(expect)
    lam2 = np.exp(1.5 * g.astype(np.float64)).ravel()
    assert np.corrcoef(counts, lam2)[0, 1] > 0.5


# ------------------------------------------------------------------ golden vectors (unmodified reference)
@pytest.mark.parametrize("case", [0, 1, 2])
def test_golden_gaussian_field(host, golden_dir, case):
    """tests/golden/ref_mock.npz: /root/reference/src/gauss_field.py:gaussian_field run UNMODIFIED on the
    uniforms our generator assigns to each mode (oracle/make_mock_golden.py)."""
    g = np.load(os.path.join(golden_dir, "ref_mock.npz"))
    n, ray, seed, box = (int(g[f"gf{case}_n"]), int(g[f"gf{case}_rayleigh"]), int(g[f"gf{case}_seed"]),
                         float(g[f"gf{case}_box"]))
    ours = host_gaussian_field(host, n, g[f"gf{case}_kf"], g[f"gf{case}_pkf"], ray, seed, box)
    ref = g[f"gf{case}_delta_k"]
    assert ref.dtype == np.complex64 and ref.shape == ours.shape
    np.testing.assert_array_equal(ours.astype(np.complex64), ref)


def test_golden_populate_field(host, golden_dir):
    g = np.load(os.path.join(golden_dir, "ref_mock.npz"))
    n, box, density, seed = int(g["pf_n"]), float(g["pf_box"]), float(g["pf_density"]), int(g["pf_seed"])
    rho = np.ascontiguousarray(g["pf_rho"])
    s = host.mock_density_sum(rho.ctypes.data, rho.size, 0, 0.0, 148 * 8, 256)
    counts = np.zeros(rho.size, dtype=np.uint32)
    total = host.mock_populate_count(rho.ctypes.data, n, box, density, 0, 0.0, seed, s, counts.ctypes.data)
    np.testing.assert_array_equal(counts, g["pf_counts"])
    pos = np.zeros((total, 3), dtype=np.float32)
    host.mock_populate_fill(counts.ctypes.data, n, np.float32(box), seed, pos.ctypes.data)
    # reference: float32 centres + float64 offsets; ours: float32 throughout (the JAX twin's arithmetic)
    d = np.abs(pos.astype(np.float64) - g["pf_coords_ref"])
    d = np.minimum(d, box - d)                                            # a wrap decided at the box edge
    assert d.max() <= 2e-7 * box
    # the JAX twin (src/populate_field.py, unmodified, run under oracle/jaxshim.py with the same draws injected)
    # is float32 throughout, like the device: bit for bit
    np.testing.assert_array_equal(pos, g["pf_coords_ref_jax"])
