"""Row f-3 (mock generator) on the GPU, through the C ABI: the kernels of csrc/mockgen.cu against the
golden vectors from the unmodified reference (tests/golden/ref_mock.npz) and, element by element,
against the host build of the same header (tests/helpers/mockgen_host.cpp), plus end-to-end sanity of
the lognormal recipe of /root/reference/tests/create_lognormal.py:44-55.  (File name sorts after the
older GPU tests on purpose: this row was added last.)"""
import os

import numpy as np
import pytest
import torch

from tests.test_mockgen_cpu import host_gaussian_field, load_host

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mocks():
    import jax_powspec_b200.mocks as m
    return m


@pytest.fixture(scope="module")
def host():
    return load_host()


def _table():
    kf = np.linspace(1e-3, 3.0, 500)
    return kf, 2.0e4 * (kf / 0.02) / (1.0 + (kf / 0.02) ** 2) ** 1.7


def _field_close(got, want):
    # same float64 arithmetic on both sides; libdevice vs libm sin/cos/log may differ in the last bit of a
    # double, i.e. at most one float32 ulp after the cast
    got, want = np.asarray(got), np.asarray(want)
    scale = np.abs(want).max()
    assert np.abs(got - want).max() <= 2.5e-7 * scale
    assert (got == want).mean() > 0.99


@pytest.mark.parametrize("case", [0, 1, 2])
def test_golden_gaussian_field(mocks, golden_dir, case):
    g = np.load(os.path.join(golden_dir, "ref_mock.npz"))
    n, ray, seed, box = (int(g[f"gf{case}_n"]), int(g[f"gf{case}_rayleigh"]), int(g[f"gf{case}_seed"]),
                         float(g[f"gf{case}_box"]))
    dk = mocks.gaussian_field(n, g[f"gf{case}_kf"], g[f"gf{case}_pkf"], ray, seed, box)
    assert dk.is_cuda and dk.dtype == torch.complex64 and tuple(dk.shape) == (n, n, n // 2 + 1)
    _field_close(dk.cpu().numpy(), g[f"gf{case}_delta_k"])


@pytest.mark.parametrize("n,rayleigh", [(32, 1), (33, 1), (48, 0)])
def test_gaussian_field_equals_host_build(mocks, host, n, rayleigh):
    kf, pkf = _table()
    box, seed = 700.0, (1 << 40) + 17                                   # 64-bit seed
    dk = mocks.gaussian_field(n, kf, pkf, rayleigh, seed, box).cpu().numpy()
    want = host_gaussian_field(host, n, kf, pkf, rayleigh, seed, box).astype(np.complex64)
    _field_close(dk, want)
    assert dk[0, 0, 0] == 0
    if n % 2 == 0:
        real = torch.fft.irfftn(torch.from_numpy(dk).cuda(), s=(n, n, n))
        back = torch.fft.rfftn(real).cpu().numpy()
        assert np.abs(back - dk).max() <= 1e-4 * np.abs(dk).max()      # Hermitian: nothing lost in the round trip


def test_gaussian_field_argument_errors(mocks):
    from jax_powspec_b200._lib import JpsError
    with pytest.raises(ValueError):
        mocks.gaussian_field(16, [0.1], [1.0], 0, 1, 100.0)
    with pytest.raises(JpsError):
        mocks.gaussian_field(16, [0.2, 0.1], [1.0, 1.0], 0, 1, 100.0)   # table not increasing
    with pytest.raises(JpsError):
        mocks.gaussian_field(16, [0.1, 0.2], [1.0, 1.0], 0, 1, -5.0)


def _host_populate(host, rho, n, box, density, seed, lognormal, bias):
    rho = np.ascontiguousarray(rho, dtype=np.float32)
    density, bias = float(np.float32(density)), float(np.float32(bias))   # the C ABI takes float32 scalars
    s = host.mock_density_sum(rho.ctypes.data, rho.size, int(lognormal), float(bias), 148 * 8, 256)
    counts = np.zeros(rho.size, dtype=np.uint32)
    total = host.mock_populate_count(rho.ctypes.data, n, box, density, int(lognormal), float(bias), seed, s,
                                     counts.ctypes.data)
    return counts.reshape(n, n, n), int(total)


@pytest.mark.parametrize("n,lognormal", [(32, False), (40, True), (17, False)])
def test_populate_equals_host_build(mocks, host, n, lognormal):
    rng = np.random.default_rng(n)
    box, density, seed = 400.0, 2.0e-3, 123456789012
    if lognormal:
        mesh = (rng.normal(size=(n, n, n)) * 0.6).astype(np.float32)
        bias = 1.4
    else:
        mesh = np.exp(rng.normal(size=(n, n, n))).astype(np.float32)
        mesh[rng.random((n, n, n)) < 0.05] = 0.0                        # empty cells
        mesh[0, 0, :4] *= 4000.0                                        # a few cells deep in the rejection sampler
        bias = None
    pos, counts = mocks.populate_field(mesh, n, box, density, seed, lognormal_bias=bias, return_counts=True)
    want_counts, total = _host_populate(host, mesh, n, box, density, seed, lognormal, bias or 0.0)
    assert isinstance(pos, np.ndarray) and pos.dtype == np.float32 and pos.shape == (counts.sum(), 3)
    # the decision boundaries of the samplers are hit with probability ~1e-15 per cell (libdevice vs libm exp / log)
    assert (counts != want_counts).sum() <= 2
    if (counts == want_counts).all():
        assert pos.shape[0] == total
        want = np.zeros((total, 3), dtype=np.float32)
        cflat = np.ascontiguousarray(want_counts.ravel())
        host.mock_populate_fill(cflat.ctypes.data, n, np.float32(box), seed, want.ctypes.data)
        d = np.abs(pos.astype(np.float64) - want)
        d = np.minimum(d, box - d)
        assert d.max() <= 2e-7 * box
        assert (pos == want).mean() > 0.999                             # same separately rounded float32 ops on both sides
    assert (pos >= 0).all() and (pos < np.float32(box)).all()
    cell = np.repeat(np.arange(n ** 3), counts.ravel())
    centre = (np.stack(np.unravel_index(cell, (n, n, n)), axis=1) + 0.5) * (box / n)
    d = np.abs(pos - centre)
    d = np.minimum(d, box - d)
    assert d.max() <= box / n * (1 + 1e-5)                              # grouped by cell in C order, within one cell


def test_populate_device_in_device_out_and_empty(mocks):
    n, box = 16, 100.0
    rho = torch.ones((n, n, n), device="cuda")
    pos = mocks.populate_field(rho, n, box, 0.05, 7)
    assert pos.is_cuda and pos.dtype == torch.float32 and pos.shape[1] == 3
    expect = 0.05 * box ** 3
    assert abs(pos.shape[0] - expect) < 5 * np.sqrt(expect)
    assert torch.equal(rho, torch.ones_like(rho))                       # the input mesh is not rescaled in place
    again = mocks.populate_field(rho, n, box, 0.05, 7)
    assert torch.equal(pos, again)                                      # same seed, same catalogue
    other = mocks.populate_field(rho, n, box, 0.05, 8)
    assert other.shape != pos.shape or not torch.equal(pos, other)
    none = mocks.populate_field(rho, n, box, 0.0, 7)
    assert none.shape == (0, 3)
    key = mocks.populate_field(rho, n, box, 0.05, np.array([0, 7], dtype=np.uint32))   # PRNGKey-like seed
    assert torch.equal(key, pos)


def test_lognormal_mock_end_to_end(mocks):
    """Gaussian field -> irfftn -> Poisson sample of exp(b g): the catalogue traces the field it was drawn from
    and carries the expected number of particles."""
    import jax_powspec_b200 as jps
    n, box, density, bias = 64, 1000.0, 3.5e-3, 1.1                      # tests/create_lognormal.py:13-19 at grid 64
    kf, pkf = _table()
    pos = mocks.lognormal_mock(n, kf, pkf, bias, density, 100, box)
    expect = density * box ** 3
    assert abs(pos.shape[0] - expect) < 6 * np.sqrt(expect)
    assert bool((pos >= 0).all()) and bool((pos < box).all())
    g = torch.fft.irfftn(mocks.gaussian_field(n, kf, pkf, 0, 100, box), s=(n, n, n))
    rho = torch.exp(bias * g)
    mesh = jps.cic_mas_vec(torch.zeros((n, n, n), device="cuda"), pos[:, 0].contiguous(), pos[:, 1].contiguous(),
                           pos[:, 2].contiguous(), None, pos.shape[0], 0.0, 0.0, 0.0, box, n, True)
    # the painted catalogue is the field smoothed by the (triangular offset * CIC) kernel plus shot noise
    a = (mesh / mesh.mean() - 1).flatten()
    b = (rho / rho.mean() - 1).flatten()
    r = float((a * b).mean() / (a.std() * b.std()))
    assert r > 0.35, r                                                   # 0.437 with the host build of the same generator
    # power on large scales: b^2 P_lin within sample variance (fixed amplitudes, a dozen modes per bin)
    kF = 2 * np.pi / box
    edges = np.arange(1.5, 8.5, 1.0) * kF
    k3d, pk, nm = jps.powspec_vec(mesh / mesh.mean() - 1, box, edges.astype(np.float32))
    pk0 = pk[:, 0].cpu().numpy() - 1.0 / density
    lin = np.interp(k3d.cpu().numpy(), kf, pkf)
    ratio = pk0 / lin
    assert 0.8 < np.median(ratio) < 1.6, ratio                           # b^2 = 1.21; host build: 1.06 .. 1.21
