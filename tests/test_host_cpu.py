"""Host-side logic of the drop-in layer that needs no GPU."""
import numpy as np


def test_triangle_pairs_order():
    import jax_powspec_b200 as jps
    k1, k2 = jps.triangle_pairs([0.1, 0.2, 0.3])
    np.testing.assert_array_equal(k1, np.float32([0.1, 0.1, 0.1, 0.2, 0.2, 0.3]))
    np.testing.assert_array_equal(k2, np.float32([0.1, 0.2, 0.3, 0.2, 0.3, 0.3]))
    assert k1.dtype == np.float32 and k2.dtype == np.float32


def test_c3_sweep_size():
    """shells centred at 2 kF j up to 0.3 h/Mpc in a 1000 Mpc/h box -> 23 centres,
    276 (k1 <= k2) pairs."""
    import jax_powspec_b200 as jps
    kF = 2 * np.pi / 1000.0
    centres = np.arange(2 * kF, 0.3, 2 * kF)
    k1, k2 = jps.triangle_pairs(centres)
    assert centres.size == 23 and k1.size == 276 and (k1 <= k2).all()
