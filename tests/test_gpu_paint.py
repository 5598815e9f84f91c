"""CUDA painters (through the C ABI) against the golden vectors and the f64 oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import mas as om
from tests.util import F32, clustered_particles

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def jps():
    import jax_powspec_b200
    return jax_powspec_b200


def _mesh_close(got, want, n_per_cell_scale=1.0):
    # float32 products + atomics in arbitrary order: error ~ few ulp of the cell value
    want = np.asarray(want, dtype=np.float64)
    tol = 4e-6 * np.maximum(np.abs(want), want.mean() * n_per_cell_scale) + 1e-7
    bad = np.abs(np.asarray(got, dtype=np.float64) - want) > tol
    assert not bad.any(), f"{bad.sum()} cells differ, max err {np.abs(got - want).max()}"


@pytest.mark.parametrize("tag", ["a", "b"])
@pytest.mark.parametrize("variant", ["vec", "scan"])
@pytest.mark.parametrize("wrap", [True, False])
@pytest.mark.parametrize("method", ["atomic", "sorted"])
def test_golden_reference_compat(jps, golden_dir, tag, variant, wrap, method):
    g = np.load(os.path.join(golden_dir, f"ref_paint_{tag}.npz"))
    p, w, n, box, xmin = g["particles"], g["weights"], int(g["n"]), float(g["box"]), float(g["xmin"])
    fn = jps.cic_mas_vec if variant == "vec" else jps.cic_mas
    got = fn(np.zeros((n, n, n), F32), p[:, 0], p[:, 1], p[:, 2], w, len(p), xmin, xmin, xmin, box, n, wrap,
             method=method)
    assert isinstance(got, np.ndarray) and got.dtype == F32
    _mesh_close(got, g[f"{variant}_wrap{int(wrap)}"])


def test_golden_accumulate_and_functional(jps, golden_dir):
    g = np.load(os.path.join(golden_dir, "ref_paint_a.npz"))
    p, w, n, box, xmin = g["particles"], g["weights"], int(g["n"]), float(g["box"]), float(g["xmin"])
    pre = torch.from_numpy(g["pre"]).cuda()
    keep = pre.clone()
    pd = torch.from_numpy(p).cuda()
    got = jps.cic_mas_vec(pre, pd[:, 0], pd[:, 1], pd[:, 2], torch.from_numpy(w).cuda(), len(p),
                          xmin, xmin, xmin, box, n, True, method="sorted")
    assert got.is_cuda and torch.equal(pre, keep), "input mesh must not be modified (functional API)"
    _mesh_close(got.cpu().numpy(), g["vec_accumulate"])


@pytest.mark.parametrize("order,compat", [(2, "reference"), (2, "fixed"), (3, "fixed"), (4, "fixed")])
@pytest.mark.parametrize("wrap", [True, False])
@pytest.mark.parametrize("method", ["atomic", "sorted"])
def test_against_f64_oracle(jps, order, compat, wrap, method):
    n, box, npart = 64, 1000.0, 300_000
    p = clustered_particles(11 + order, npart, box)
    w = (0.5 + np.random.default_rng(3).random(npart)).astype(F32)
    want = om.paint(np.zeros((n, n, n)), p[:, 0], p[:, 1], p[:, 2], w, 0.0, 0.0, 0.0, box, n, wrap,
                    order=order, compat=compat, precision="f64")
    got = jps.paint(np.zeros((n, n, n), F32), p[:, 0], p[:, 1], p[:, 2], w, 0.0, 0.0, 0.0, box, n, wrap,
                    order=order, compat=compat, method=method)
    _mesh_close(got, want)
    if compat == "fixed" and wrap:
        assert abs(got.sum(dtype=np.float64) - w.sum(dtype=np.float64)) < 2e-6 * w.sum()   # mass conservation


@pytest.mark.parametrize("method", ["atomic", "sorted"])
def test_unweighted_strided_columns_nonpow2(jps, method):
    n, box, npart = 50, 600.0, 100_000          # non power of two mesh (tests/void-model.py:69 uses 300)
    p = clustered_particles(5, npart, box)
    pd = torch.from_numpy(p).cuda()             # (Np,3) rows: x,y,z are stride-3 views, no copy
    got = jps.paint(torch.zeros((n, n, n), device="cuda"), pd[:, 0], pd[:, 1], pd[:, 2], None,
                    0.0, 0.0, 0.0, box, n, True, order=2, compat="reference", method=method)
    want = om.paint(np.zeros((n, n, n)), p[:, 0], p[:, 1], p[:, 2], None, 0.0, 0.0, 0.0, box, n, True,
                    order=2, compat="reference", precision="f64")
    _mesh_close(got.cpu().numpy(), want)


def test_known_answers(jps):
    n, box = 8, 8.0
    z = np.zeros((n, n, n), F32)
    one = np.ones(1, F32)
    # particle on a node -> one cell
    m = jps.cic_mas_vec(z, np.array([3.0], F32), np.array([2.0], F32), np.array([5.0], F32), one, 1, 0., 0., 0., box, n, True)
    assert m[3, 2, 5] == 1.0 and m.sum() == 1.0
    # mid-cell: textbook 8 x 1/8; the reference's Q1 corner gets mdx*mdy*ddz = 1/8 as well at d=1/2
    m = jps.cic_mas_vec(z, np.array([3.5], F32), np.array([2.5], F32), np.array([7.5], F32), one, 1, 0., 0., 0., box, n, True)
    assert np.isclose(m.sum(), 1.0) and np.isclose(m[3, 2, 7], 0.125) and np.isclose(m[4, 3, 0], 0.125)  # z wraps
    # Q1: off-centre particle, mass = 1 + mdx*ddz*(mdy-ddy)
    dx, dy, dz = 0.25, 0.125, 0.75
    m = jps.cic_mas_vec(z, np.array([1 + dx], F32), np.array([1 + dy], F32), np.array([1 + dz], F32), one, 1, 0., 0., 0., box, n, True)
    assert np.isclose(m.sum(), 1 + (1 - dx) * dz * ((1 - dy) - dy), atol=1e-6)
    m = jps.cic_mas_vec(z, np.array([1 + dx], F32), np.array([1 + dy], F32), np.array([1 + dz], F32), one, 1, 0., 0., 0., box, n, True, compat="fixed")
    assert np.isclose(m.sum(), 1.0, atol=1e-6)
    # TSC on a node: 3/4, 1/8, 1/8 per axis;  PCS on a node: 2/3, 1/6, 1/6
    m = jps.tsc_mas_vec(z, np.array([4.0], F32), np.array([4.0], F32), np.array([4.0], F32), one, 1, 0., 0., 0., box, n, True)
    assert np.isclose(m[4, 4, 4], 0.75 ** 3) and np.isclose(m[3, 4, 4], 0.125 * 0.75 ** 2) and np.isclose(m.sum(), 1.0)
    m = jps.pcs_mas_vec(z, np.array([4.0], F32), np.array([4.0], F32), np.array([4.0], F32), one, 1, 0., 0., 0., box, n, True)
    assert np.isclose(m[4, 4, 4], (2 / 3) ** 3) and np.isclose(m[5, 4, 4], (1 / 6) * (2 / 3) ** 2) and np.isclose(m.sum(), 1.0)
    # empty input
    m = jps.cic_mas_vec(z, np.zeros(0, F32), np.zeros(0, F32), np.zeros(0, F32), np.zeros(0, F32), 0, 0., 0., 0., box, n, True)
    assert m.sum() == 0.0


@pytest.mark.parametrize("order", [2, 3, 4])
def test_translation_by_whole_cells(jps, order):
    n, box, npart = 32, 320.0, 50_000
    p = clustered_particles(21, npart, box)
    cell = box / n
    base = jps.paint(np.zeros((n, n, n), F32), p[:, 0], p[:, 1], p[:, 2], None, 0., 0., 0., box, n, True,
                     order=order, compat="fixed", method="atomic")
    # shifting the origin by whole cells rolls the mesh (xmin is per axis)
    shifted = jps.paint(np.zeros((n, n, n), F32), p[:, 0], p[:, 1], p[:, 2], None, -3 * cell, 2 * cell, 0., box, n, True,
                        order=order, compat="fixed", method="atomic")
    np.testing.assert_allclose(shifted, np.roll(base, (3, -2), axis=(0, 1)), rtol=2e-4, atol=2e-4)


def test_bad_arguments(jps):
    z = np.zeros((8, 8, 8), F32)
    a = np.zeros(4, F32)
    with pytest.raises(ValueError):
        jps.cic_mas_vec(z, a, a, a, a, 4, 0., 0., 0., 8.0, 16, True)        # n_bins != mesh shape
    with pytest.raises(jps._lib.JpsError):
        jps.paint(z, a, a, a, a, 0., 0., 0., 8.0, 8, True, order=5)
    with pytest.raises(jps._lib.JpsError):
        jps.paint(z, a, a, a, a, 0., 0., 0., -1.0, 8, True)


@pytest.mark.parametrize("order", [2, 3, 4])
def test_sorted_equals_atomic_large_mesh_properties(jps, order):
    """BASELINE-sized mesh (512^3, bigger than L2), 2e7 clustered particles: the bucketed painter
    against the per-particle one, and mass conservation -- size-independent properties."""
    from jax_powspec_b200.mocks import lognormal_catalog
    n, box, npart = 512, 2000.0, 20_000_000
    x, y, z = lognormal_catalog(npart, box, n_grid=128, seed=9, device="cuda")
    zero = torch.zeros((n, n, n), device="cuda")
    a = jps.paint(zero, x, y, z, None, 0., 0., 0., box, n, True, order=order, compat="fixed", method="atomic")
    b = jps.paint(zero, x, y, z, None, 0., 0., 0., box, n, True, order=order, compat="fixed", method="sorted")
    assert abs(b.sum(dtype=torch.float64).item() - npart) < 2e-6 * npart
    diff = (a - b).abs()
    tol = 2e-5 * torch.maximum(a.abs(), torch.tensor(1.0, device="cuda")) + 1e-6      # two float32 summation orders
    assert bool((diff <= tol).all()), f"max diff {diff.max().item()}"


@pytest.mark.parametrize("order,compat", [(2, "reference"), (3, "fixed"), (4, "fixed")])
@pytest.mark.parametrize("method", ["atomic", "sorted"])
def test_float32_grid_position_is_not_fused(jps, order, compat, method):
    """Regression: pos = (x-xmin)*inv must be rounded to float32 BEFORE the in-cell offset is
    taken (the reference materialises it, Q4).  An FMA-contracted pos - floor(pos) shifts weights by
    up to half an ulp of pos -- 1.5e-5 relative at grid coordinate ~250 -- so particles are placed
    in the top planes of a 256^3 mesh where that is far above the 4e-6 tolerance."""
    n, box, npart = 256, 2500.0, 400_000
    rng = np.random.default_rng(123)
    p = rng.random((npart, 3)).astype(F32)
    p = (p * np.array([0.06, 1.0, 1.0], dtype=F32) + np.array([0.94, 0.0, 0.0], dtype=F32)) * F32(box)
    p[p >= F32(box)] = 0.0
    xmin = 0.0
    want = om.paint(np.zeros((n, n, n)), p[:, 0], p[:, 1], p[:, 2], None, xmin, xmin, xmin, box, n, True,
                    order=order, compat=compat, precision="f64")
    got = jps.paint(np.zeros((n, n, n), F32), p[:, 0], p[:, 1], p[:, 2], None, xmin, xmin, xmin, box, n, True,
                    order=order, compat=compat, method=method)
    err = np.abs(got - want) / np.maximum(np.abs(want), 1.0)
    assert err.max() < 2e-6, f"max error relative to max(|cell|,1): {err.max():.3e}"


def test_two_level_partition_in_subprocess():
    """The bucketing flavour is chosen once per process (JPS_BUCKET); run the two-level partition in its own."""
    import subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, JPS_BUCKET="two")
    r = subprocess.run([sys.executable, os.path.join(root, "tests", "helpers", "two_level_check.py")],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "two-level ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("cfg", [{"JPS_FINE": "staged"}, {"JPS_FINE": "direct"}, {"JPS_FINE": "staged", "JPS_FINE_CHUNK": "big"},
                                 {"JPS_TILE_ORDER": "bank"}, {"JPS_TILE_ORDER": "bank", "JPS_TILE_FLUSH": "red", "JPS_FX_BITS": "28"}])
def test_fine_pass_forms_in_subprocess(cfg):
    """Both forms of the fine pass (lone stores / chunks staged in shared memory, both chunk shapes), with few groups
    so that a group holds many tiles even on a small mesh (64 groups: 128 tiles per group at 256^3); and the deposit
    with the particles of a tile in bank-class order (an option: the default is arrival order)."""
    import subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, JPS_BUCKET="two", JPS_MAX_GROUPS="64", **cfg)
    r = subprocess.run([sys.executable, os.path.join(root, "tests", "helpers", "two_level_check.py")],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "two-level ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("cfg", [{"JPS_COUNT": "fused", "JPS_FINE": "staged", "JPS_FINE_CHUNK": "big"},
                                 {"JPS_COUNT": "separate", "JPS_FINE": "staged", "JPS_FINE_CHUNK": "small"},
                                 {"JPS_COUNT": "fused", "JPS_FINE": "direct"}, {}])
def test_big_mesh_bucketing_options_in_subprocess(cfg):
    """N = 592 (more tiles than the shared-memory histogram holds): tile histogram from its own pass or fused into the
    coarse pass, both forms of the fine pass, both chunk shapes -- each against the plain atomic painter."""
    import subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tests", "helpers", "big_mesh_check.py")],
                       env=dict(os.environ, **cfg), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "big-mesh ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("order", [2, 3, 4])
def test_garbage_positions_do_not_crash_or_corrupt(jps, order):
    """NaN / inf / absurd coordinates must not fault or write out of bounds (nothing is validated, as
    in the reference); in the bucketed painter a non-finite contribution rounds to zero, so the mesh
    equals the mesh of the well-formed particles."""
    n, box, npart = 128, 1000.0, 300_000
    p = clustered_particles(77, npart, box)
    bad = np.array([[np.nan, 1.0, 2.0], [np.inf, 5.0, 5.0], [-np.inf, 5.0, 5.0], [1e30, -1e30, 3.0],
                    [3.0, np.nan, np.nan], [-1e9, 2e9, 7.0]], dtype=F32)
    q = np.concatenate([p, bad])
    zero = np.zeros((n, n, n), F32)
    good = jps.paint(zero, p[:, 0], p[:, 1], p[:, 2], None, 0., 0., 0., box, n, True, order=order, compat="fixed", method="sorted")
    got = jps.paint(zero, q[:, 0], q[:, 1], q[:, 2], None, 0., 0., 0., box, n, True, order=order, compat="fixed", method="sorted")
    assert np.isfinite(got).all()
    # the finite-but-absurd ones wrap periodically and land somewhere; NaN/inf ones vanish
    assert abs(float(got.sum(dtype=np.float64)) - float(good.sum(dtype=np.float64))) <= len(bad) + 1e-3
    jps.paint(zero, q[:, 0], q[:, 1], q[:, 2], None, 0., 0., 0., box, n, True, order=order, compat="fixed", method="atomic")
    import torch
    torch.cuda.synchronize()          # no sticky CUDA error
    jps.cic_mas_vec(zero, q[:, 0], q[:, 1], q[:, 2], np.ones(len(q), F32), len(q), 0., 0., 0., box, n, True, method="sorted")
    torch.cuda.synchronize()


@pytest.mark.parametrize("order,compat", [(2, "reference"), (3, "fixed"), (4, "fixed")])
@pytest.mark.parametrize("wkind", ["mixed_sign", "wide_range", "tiny", "huge", "zeros"])
def test_fixed_point_deposit_weight_ranges(jps, order, compat, wkind):
    """The bucketed painter accumulates tiles in 64-bit fixed point scaled by 2^e >= max|w|: negative
    weights (borrow path), 6 decades of dynamic range, very small / very large scales and all-zero
    weights must all agree with the f64 oracle to float32 accuracy RELATIVE TO max|w|."""
    n, box, npart = 64, 1000.0, 200_000
    p = clustered_particles(9, npart, box)
    rng = np.random.default_rng(17)
    w = {"mixed_sign": rng.uniform(-1, 1, npart),
         "wide_range": 10.0 ** rng.uniform(-3, 3, npart) * rng.choice([-1.0, 1.0], npart),
         "tiny": rng.uniform(0.5, 1, npart) * 1e-20,
         "huge": rng.uniform(0.5, 1, npart) * 1e20,
         "zeros": np.zeros(npart)}[wkind].astype(F32)
    want = om.paint(np.zeros((n, n, n)), p[:, 0], p[:, 1], p[:, 2], w, 0., 0., 0., box, n, True,
                    order=order, compat=compat, precision="f64")
    got = jps.paint(np.zeros((n, n, n), F32), p[:, 0], p[:, 1], p[:, 2], w, 0., 0., 0., box, n, True,
                    order=order, compat=compat, method="sorted").astype(np.float64)
    wmax = float(np.abs(w).max())
    if wmax == 0.0:
        assert np.all(got == 0.0)
        return
    # absolute accuracy: float32 rounding of every product (6e-8 * |contribution|) plus 2^-31 wmax per update
    per_cell = om.paint(np.zeros((n, n, n)), p[:, 0], p[:, 1], p[:, 2], np.abs(w), 0., 0., 0., box, n, True,
                        order=order, compat=compat, precision="f64")          # sum of |contributions|
    tol = 3e-7 * per_cell + 1e-7 * wmax
    bad = np.abs(got - want) > tol
    assert not bad.any(), f"{bad.sum()} cells off, worst {np.max(np.abs(got - want) / (tol + 1e-300)):.2f}x tol"


@pytest.mark.parametrize("order", [2, 3, 4])
def test_one_outlier_weight_costs_precision_in_its_own_tile_only(jps, order):
    """The fixed-point scale is per TILE (2^e >= the tile's own max|w|): a particle 1e9 times heavier than the
    rest must not degrade cells of other tiles (with a mesh-global scale every unit-weight contribution would
    be quantised at 2^-31 * 1e9 = 0.5)."""
    n, box, npart = 64, 1000.0, 100_000
    p = clustered_particles(21, npart, box)
    w = np.ones(npart, F32)
    p[0] = [383.0, 383.0, 383.0]                    # cell 24 of every axis: anchors 23 / 24 -> tile (1, 1, 1) for every order
    w[0] = 1e9
    want = om.paint(np.zeros((n, n, n)), p[:, 0], p[:, 1], p[:, 2], w, 0., 0., 0., box, n, True,
                    order=order, compat="fixed", precision="f64")
    got = jps.paint(np.zeros((n, n, n), F32), p[:, 0], p[:, 1], p[:, 2], w, 0., 0., 0., box, n, True,
                    order=order, compat="fixed", method="sorted").astype(np.float64)
    far = np.ones((n, n, n), bool)
    far[16:36, 16:36, 16:36] = False                # tile (1,1,1) + halo: only float32 accuracy relative to 1e9 there
    err = np.abs(got - want)[far] / np.maximum(want[far], 1.0)
    assert err.max() <= 1e-6, f"cells far from the outlier are off by {err.max():.2e}"
    near = ~far
    assert np.abs(got - want)[near].max() <= 3e-7 * 1e9


@pytest.mark.parametrize("order,compat", [(2, "reference"), (2, "fixed"), (3, "fixed"), (4, "fixed")])
def test_heavy_tiles_are_split_over_ctas(jps, order, compat):
    """One CTA deposits at most 16384 particles of a tile; the rest of a heavy tile is taken by further CTAs, each
    with its own shared-memory copy of the tile (the flush is additive).  60 % of the catalogue sits in a blob narrower
    than a tile (~1.7e5 particles in one tile: 11 parts) and 20 % in a blob around the box corner (periodic wrap,
    boundary tiles that flush with per-thread reds): the mesh must agree with the f64 oracle like any other."""
    n, box, npart = 64, 1000.0, 300_000
    rng = np.random.default_rng(5)
    p = rng.random((npart, 3)) * box
    p[:180_000] = np.array([625.0, 635.0, 645.0]) + rng.normal(size=(180_000, 3)) * 20.0      # sigma = 1.3 cells, mid-tile
    p[180_000:240_000] = rng.normal(size=(60_000, 3)) * 15.0                                   # around the origin
    p = (p % box).astype(F32)
    p[p >= F32(box)] = 0.0
    w = rng.uniform(-0.5, 1.5, npart).astype(F32)
    tiles = (np.floor(p * F32(n / box)).astype(np.int64) % n) // 16
    assert np.bincount((tiles[:, 0] * 4 + tiles[:, 1]) * 4 + tiles[:, 2]).max() > 5 * 16384
    for wt in (None, w):
        want = om.paint(np.zeros((n, n, n)), p[:, 0], p[:, 1], p[:, 2], wt, 0., 0., 0., box, n, True,
                        order=order, compat=compat, precision="f64")
        per_cell = want if wt is None else om.paint(np.zeros((n, n, n)), p[:, 0], p[:, 1], p[:, 2], np.abs(wt), 0., 0., 0.,
                                                    box, n, True, order=order, compat=compat, precision="f64")
        got = jps.paint(np.zeros((n, n, n), F32), p[:, 0], p[:, 1], p[:, 2], wt, 0., 0., 0., box, n, True,
                        order=order, compat=compat, method="sorted").astype(np.float64)
        tol = 3e-7 * np.abs(per_cell) + 2e-7
        bad = np.abs(got - want) > tol
        assert not bad.any(), f"{bad.sum()} cells off, worst {np.max(np.abs(got - want) / tol):.2f}x tol (weights: {wt is not None})"
        assert abs(got.sum() - want.sum()) <= 1e-6 * np.abs(per_cell).sum()


def test_heavy_group_is_sliced_over_ctas(jps):
    """The fine pass gives a group of tiles to one CTA; a group above 2^22 records is left to fine_heavy_kernel, which
    cuts it into slices placed through the global tile cursors.  5e6 of 6e6 particles sit in one tile of a 64^3 mesh
    (one tile per group there): mesh against the float64 C oracle, and mass conservation."""
    from oracle import cport
    n, box, npart, nb = 64, 1000.0, 6_000_000, 5_000_000
    rng = np.random.default_rng(8)
    p = rng.random((npart, 3)) * box
    p[:nb] = np.array([625.0, 635.0, 645.0]) + rng.normal(size=(nb, 3)) * 20.0
    p = (p % box).astype(F32)
    p[p >= F32(box)] = 0.0
    rng.shuffle(p, axis=0)
    x, y, z = (np.ascontiguousarray(p[:, i]) for i in range(3))
    for order in (3, 4):
        want = cport.paint_f64(np.zeros((n, n, n)), x, y, z, None, 0., 0., 0., box, n, True, order=order, compat="fixed")
        got = jps.paint(np.zeros((n, n, n), F32), x, y, z, None, 0., 0., 0., box, n, True, order=order, compat="fixed",
                        method="sorted").astype(np.float64)
        err = np.abs(got - want) / np.maximum(np.abs(want), 1.0)
        # ~300 parts of the heavy tile are added to the mesh in float32 (reduce-add in L2, any order)
        assert err.max() < 3e-6, f"order {order}: max error relative to max(|cell|,1): {err.max():.3e}"
        assert abs(got.sum() / npart - 1.0) < 1e-7


@pytest.mark.parametrize("bad", [np.inf, -np.inf, np.nan])
def test_non_finite_weight_propagates_like_a_float_scatter(jps, bad):
    """inf / NaN weights never enter the fixed-point tile: the particle is deposited with float atomics, so its
    stencil cells become non-finite (as in the reference's float32 scatter) and every other cell is exact."""
    n, box, npart = 64, 1000.0, 50_000
    p = clustered_particles(22, npart, box)
    w = np.ones(npart, F32)
    p[7] = [500.0, 500.0, 500.0]
    w[7] = bad
    got = jps.paint(np.zeros((n, n, n), F32), p[:, 0], p[:, 1], p[:, 2], w, 0., 0., 0., box, n, True,
                    order=3, compat="fixed", method="sorted")
    ref = jps.paint(np.zeros((n, n, n), F32), p[:, 0], p[:, 1], p[:, 2], w, 0., 0., 0., box, n, True,
                    order=3, compat="fixed", method="atomic")
    assert np.array_equal(np.isfinite(got), np.isfinite(ref))
    assert 1 <= (~np.isfinite(got)).sum() <= 27
    ok = np.isfinite(ref)
    np.testing.assert_allclose(got[ok], ref[ok], rtol=0, atol=4e-6 * max(ref[ok].max(), 1.0))
