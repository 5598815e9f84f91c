"""Slab-sharded path: P virtual ranks driven on ONE GPU (exchanges done with tensor copies) must
reproduce the single-GPU pipeline and the f64 oracle.  The real multi-process run over NCCL is
tools/slab_check.py (gpurun --gpus 2/8); the exchange choreography itself is covered on CPU with gloo
(tests/test_dist_cpu.py)."""
import numpy as np
import pytest
import torch

from oracle import correlations as oc
from oracle import mas as om
from tests.util import F32, clustered_particles, rel_to_monopole

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def jps():
    import jax_powspec_b200
    return jax_powspec_b200


def _split_by_slab(p, box, n, world):
    owner = (np.floor(p[:, 0] * np.float32(n / box)).astype(np.int64) % n) // (n // world)
    out = []
    for r in range(world):
        q = p[owner == r]
        out.append(tuple(torch.from_numpy(np.ascontiguousarray(q[:, i])).cuda() for i in range(3)) + (None,))
    return out


@pytest.mark.parametrize("world", [1, 2, 4])
@pytest.mark.parametrize("order", [2, 3, 4])
@pytest.mark.parametrize("method", ["atomic", "sorted"])
def test_virtual_ranks_match_single_gpu_and_oracle(jps, world, order, method):
    from jax_powspec_b200.slab import SlabPipeline, run_virtual_ranks
    n, box, npart = 64, 1000.0, 300_000
    p = clustered_particles(100 + order, npart, box)
    kF = 2 * np.pi / box
    ke = np.arange(kF, np.pi * n / box, kF).astype(F32)
    pipes = [SlabPipeline(n, box, ke, order=order, compat="fixed", method=method, rank=r, world=world)
             for r in range(world)]
    outs = run_virtual_ranks(pipes, _split_by_slab(p, box, n, world))
    k3d, pk, nm = (t.cpu().numpy() for t in outs[0])
    for o in outs[1:]:                                   # every rank ends with the same answer
        np.testing.assert_array_equal(o[1].cpu().numpy(), pk)
    # the painted slabs tile the single-GPU mesh
    ref_mesh = jps.paint(np.zeros((n, n, n), F32), p[:, 0], p[:, 1], p[:, 2], None, 0., 0., 0., box, n, True,
                         order=order, compat="fixed", method="atomic")
    got_mesh = np.concatenate([q.owned().cpu().numpy() for q in pipes], axis=0)
    np.testing.assert_allclose(got_mesh, ref_mesh, rtol=2e-5, atol=2e-5)
    # multipoles: single-GPU fused pipeline and f64 oracle
    k1, pk1, nm1 = jps.paint_powspec(p[:, 0], p[:, 1], p[:, 2], None, 0., 0., 0., box, n, ke, order=order,
                                     compat="fixed", method="atomic")
    np.testing.assert_array_equal(nm, nm1)
    np.testing.assert_array_equal(k3d, k1)
    assert rel_to_monopole(pk.astype(np.float64), pk1.astype(np.float64)).max() < 1e-5
    rho = om.paint(np.zeros((n, n, n)), p[:, 0], p[:, 1], p[:, 2], None, 0., 0., 0., box, n, True,
                   order=order, compat="fixed", precision="f64")
    delta = (rho / rho.mean() - 1.0).astype(F32)
    _, pk64, counts = oc.powspec(delta, box, ke, mas_order=order, precision="f64")
    np.testing.assert_array_equal(nm.astype(np.int64), counts)
    assert rel_to_monopole(pk.astype(np.float64), pk64).max() < 2e-5
    for q in pipes:
        q.close()


@pytest.mark.parametrize("layout", ["xslow", "xfast", "xfast-pencil"])
@pytest.mark.parametrize("world,n", [(2, 64), (4, 64), (2, 96), (8, 64)])
def test_fused_peer_store_transpose_both_layouts(jps, world, n, layout):
    """The p2p transport on one device (peers = the other virtual ranks' receive buffers): the plain
    peer-store kernel and the transposing one (x-fast shard, contiguous 1-D FFT, kx-lane binning)
    must reproduce the tensor-copy exchange bit for bit in the counts and to rounding in P."""
    from jax_powspec_b200.slab import SlabPipeline, run_virtual_ranks
    box, npart, order = 1000.0, 200_000, 3
    p = clustered_particles(300 + world, npart, box)
    kF = 2 * np.pi / box
    ke = np.arange(kF, np.pi * n / box, kF).astype(F32)
    cats = _split_by_slab(p, box, n, world)
    ref_pipes = [SlabPipeline(n, box, ke, order=order, compat="fixed", rank=r, world=world) for r in range(world)]
    k0, pk0, nm0 = (t.cpu().numpy() for t in run_virtual_ranks(ref_pipes, cats)[0])
    fft = "pencil" if layout == "xfast-pencil" else "cufft2d"      # pencil: C2C n/2 + fused untangle/transpose + C2C along y
    layout = layout.split("-")[0]
    pipes = [SlabPipeline(n, box, ke, order=order, compat="fixed", rank=r, world=world, fft=fft) for r in range(world)]
    for q in pipes:
        q._force_chunks = (n == 96)                      # also the chunked FFT / transfer overlap path
    outs = run_virtual_ranks(pipes, cats, p2p=layout)
    assert all(q.transport == "p2p" and q.xfast == (layout == "xfast") for q in pipes)
    k1, pk1, nm1 = (t.cpu().numpy() for t in outs[0])
    np.testing.assert_array_equal(nm1, nm0)
    np.testing.assert_array_equal(k1, k0)
    assert rel_to_monopole(pk1.astype(np.float64), pk0.astype(np.float64)).max() < 2e-6
    for q in pipes + ref_pipes:
        q.close()


def test_single_rank_pipeline_call(jps):
    """world_size 1 through SlabPipeline.__call__ (no process group): same code path the multi-GPU run takes."""
    from jax_powspec_b200.slab import SlabPipeline
    n, box, npart = 128, 2000.0, 1_000_000
    p = clustered_particles(5, npart, box)
    ke = np.arange(0.01, 0.19, 0.005).astype(F32)
    pipe = SlabPipeline(n, box, ke, order=3)
    x, y, z = (torch.from_numpy(np.ascontiguousarray(p[:, i])).cuda() for i in range(3))
    k3d, pk, nm = pipe(x, y, z)
    k1, pk1, nm1 = jps.paint_powspec(x, y, z, None, 0., 0., 0., box, n, ke, order=3, compat="fixed")
    assert torch.equal(nm, nm1)
    assert rel_to_monopole(pk.cpu().numpy().astype(np.float64), pk1.cpu().numpy().astype(np.float64)).max() < 1e-5
