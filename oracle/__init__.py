"""CPU oracle for the jax-powspec hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import anything from this package; the product
(``jax_powspec_b200``) never does and fails loudly when its CUDA library is
missing.

Parity status: the reference ships no golden vectors, no assertions and no input
data and JAX is not installable here, so there is no
output of the *real* JAX reference to pin against.  The strongest pin available
is used instead: the UNMODIFIED reference sources are executed on NumPy through
``oracle/jaxshim.py`` (a stand-in for the JAX API subset they use) by
``oracle/run_reference.py``; its outputs are committed under ``tests/golden/``
and the restatement in ``oracle/mas.py`` / ``oracle/correlations.py`` is checked
against them.  TSC / PCS painting and their window corrections do not exist in
the reference at all: for those rows parity is UNPINNED (analytic known-answer
tests only).
"""
