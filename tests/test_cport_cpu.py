"""The C restatement (oracle/c/jps_oracle.c) -- the CPU baseline bench.py reports and the oracle the
large-size tools use -- against the golden vectors of the shim-run reference and against the NumPy
restatement.  CPU only."""
import os

import numpy as np
import pytest

from oracle import correlations as oc
from oracle import cport
from oracle import mas as om

F32 = np.float32


@pytest.mark.parametrize("tag", ["a", "b"])
@pytest.mark.parametrize("variant", ["vec", "scan"])
@pytest.mark.parametrize("wrap", [True, False])
def test_c_painter_matches_the_reference_bit_for_bit(golden_dir, tag, variant, wrap):
    g = np.load(os.path.join(golden_dir, f"ref_paint_{tag}.npz"))
    p, w, n, box, xmin = g["particles"], g["weights"], int(g["n"]), float(g["box"]), float(g["xmin"])
    got = cport.paint(np.zeros((n, n, n), F32), p[:, 0], p[:, 1], p[:, 2], w, xmin, xmin, xmin, box, n, wrap,
                      order=2, compat="reference", variant=variant)
    np.testing.assert_array_equal(got, g[f"{variant}_wrap{int(wrap)}"])     # same serial float32 order as XLA CPU


def test_c_painter_accumulates(golden_dir):
    g = np.load(os.path.join(golden_dir, "ref_paint_a.npz"))
    p, w, n, box, xmin = g["particles"], g["weights"], int(g["n"]), float(g["box"]), float(g["xmin"])
    got = cport.paint(g["pre"], p[:, 0], p[:, 1], p[:, 2], w, xmin, xmin, xmin, box, n, True)
    np.testing.assert_array_equal(got, g["vec_accumulate"])


@pytest.mark.parametrize("order", [2, 3, 4])
@pytest.mark.parametrize("weighted", [False, True])
def test_c_bspline_painter_matches_numpy_oracle(order, weighted):
    rng = np.random.default_rng(order)
    n, box = 20, 500.0
    p = (rng.random((4000, 3)) * box).astype(F32)
    p[p >= F32(box)] = 0.0
    w = (rng.random(4000).astype(F32) + F32(0.5)) if weighted else None
    got = cport.paint(np.zeros((n, n, n), F32), p[:, 0], p[:, 1], p[:, 2], w, 0.0, 0.0, 0.0, box, n, True,
                      order=order, compat="fixed")
    want = om.paint(np.zeros((n, n, n)), p[:, 0], p[:, 1], p[:, 2], w, 0.0, 0.0, 0.0, box, n, True,
                    order=order, compat="fixed", precision="f64")
    np.testing.assert_allclose(got, want, rtol=0, atol=3e-6 * want.max())
    total = 4000.0 if w is None else float(w.astype(np.float64).sum())
    assert got.astype(np.float64).sum() == pytest.approx(total, rel=1e-5)


@pytest.mark.parametrize("tag", ["a", "b", "c"])
@pytest.mark.parametrize("edges", ["kf", "fine", "wide"])
def test_c_binning_matches_the_reference(golden_dir, tag, edges):
    g = np.load(os.path.join(golden_dir, f"ref_corr_{tag}.npz"))
    k3d, pk, nm = cport.powspec(g["delta"], float(g["box"]), g[f"pk_{edges}_edges"])
    want_pk, want_nm = g[f"pk_{edges}_Pk3D"], g[f"pk_{edges}_Nmodes3D"]
    np.testing.assert_array_equal(nm, want_nm)                                # float32 counts, exact
    np.testing.assert_array_equal(k3d, g[f"pk_{edges}_k3D"])
    ok = want_nm > 0
    assert np.all(np.isnan(pk[~ok]))
    scale = np.abs(want_pk[ok][:, :1])
    np.testing.assert_allclose(pk[ok] / scale, want_pk[ok] / scale, rtol=0, atol=2e-5)


@pytest.mark.parametrize("mas_order", [3, 4])
def test_c_binning_other_windows(mas_order):
    rng = np.random.default_rng(9)
    n, box = 24, 300.0
    delta = rng.normal(size=(n, n, n)).astype(F32)
    kF = 2 * np.pi / box
    ke = np.arange(kF, np.pi * n / box, kF).astype(F32)
    _, pk, nm = cport.powspec(delta, box, ke, mas_order=mas_order)
    _, pk64, counts = oc.powspec(delta, box, ke, mas_order=mas_order, precision="f64")
    np.testing.assert_array_equal(nm.astype(np.int64), counts)
    np.testing.assert_allclose(pk / np.abs(pk64[:, :1]), pk64 / np.abs(pk64[:, :1]), rtol=0, atol=2e-5)


# ---- the float64 C + OpenMP restatement used by the FULL-SIZE GPU parity tests (tests/test_gpu_fullsize.py):
#      pinned here on the NumPy float64 oracle, which is itself pinned on the shim-run reference
@pytest.mark.parametrize("order,compat", [(2, "reference"), (2, "fixed"), (3, "fixed"), (4, "fixed")])
@pytest.mark.parametrize("wrap", [True, False])
def test_c_f64_painter_equals_numpy_f64_oracle(order, compat, wrap):
    rng = np.random.default_rng(10 * order + wrap)
    n, box = 24, 500.0
    p = (rng.random((20000, 3)) * box * 1.04 - 0.02 * box).astype(F32)      # a few particles outside the box
    if compat == "fixed" or wrap:
        pass
    w = rng.random(20000).astype(F32) + F32(0.5)
    pre = rng.random((n, n, n))
    for variant in (("vec", "scan") if compat == "reference" else ("vec",)):
        got = cport.paint_f64(pre, p[:, 0], p[:, 1], p[:, 2], w, -1.5, 0.25, 0.0, box, n, wrap, order=order,
                              compat=compat, variant=variant)
        want = om.paint(pre, p[:, 0], p[:, 1], p[:, 2], w, -1.5, 0.25, 0.0, box, n, wrap, order=order,
                        compat=compat, variant=variant, precision="f64")
        np.testing.assert_allclose(got, want, rtol=0, atol=1e-12 * want.max())


@pytest.mark.parametrize("mas_order", [2, 3, 4])
def test_c_f64_binning_equals_numpy_f64_oracle(mas_order):
    rng = np.random.default_rng(mas_order)
    n, box = 36, 700.0
    delta = rng.normal(size=(n, n, n)).astype(F32)
    for ke in (np.arange(2 * np.pi / box, np.pi * n / box, 2 * np.pi / box).astype(F32),
               np.arange(1e-4, 5, 0.2e-2).astype(F32)):
        k3d, pk, cnt = cport.powspec_f64(delta, box, ke, mas_order=mas_order)
        k64, pk64, c64 = oc.powspec(delta, box, ke, mas_order=mas_order, precision="f64")
        np.testing.assert_array_equal(cnt, c64)
        np.testing.assert_array_equal(k3d, k64)
        ok = c64 > 0
        np.testing.assert_allclose(pk[ok], pk64[ok], rtol=1e-11, atol=0)


def test_c_multicore_painter_equals_serial_port():
    rng = np.random.default_rng(3)
    n, box = 32, 100.0
    p = (rng.random((50000, 3)) * box).astype(F32)
    p[p >= F32(box)] = 0.0
    for order in (2, 3, 4):
        a = cport.paint(np.zeros((n, n, n), F32), p[:, 0], p[:, 1], p[:, 2], None, 0., 0., 0., box, n, True,
                        order=order, compat="fixed")
        b = cport.paint_mt(np.zeros((n, n, n), F32), p[:, 0], p[:, 1], p[:, 2], None, 0., 0., 0., box, n, True, order=order)
        np.testing.assert_allclose(a, b, rtol=0, atol=2e-5 * a.max())
