"""world_size-2 gloo tests (CPU) of the multi-process host logic: realisation sharding, gather,
max-over-ranks and the covariance batch driver.  The compute callback is a stand-in; the GPU
pipeline itself is covered by the -m gpu tests."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_measure(seed):
    g = torch.Generator().manual_seed(int(seed))
    return torch.rand((5, 3), generator=g)


def _worker(rank, world_size, port, n_items, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world_size), LOCAL_RANK=str(rank))
    from jax_powspec_b200 import dist as jd
    r, w, _ = jd.init_from_env(backend="gloo")
    assert (r, w) == (rank, world_size) == jd.world()
    mine = jd.shard_indices(n_items)
    assert mine == list(range(rank, n_items, world_size))
    assert jd.max_over_ranks(10.0 + rank) == 10.0 + (world_size - 1)
    local = torch.tensor([[float(i), float(i) * 2] for i in mine], dtype=torch.float32).reshape(len(mine), 2)
    full = jd.gather_rows(local, n_items)
    if rank == 0:
        assert torch.equal(full, torch.tensor([[float(i), float(i) * 2] for i in range(n_items)]))
    else:
        assert full is None
    # generic sharded units (the bispectrum sweep's host logic): tuple of row tensors, one call per rank
    calls = []

    def compute(idx):
        calls.append(list(idx))
        i = torch.tensor(idx, dtype=torch.float32)
        return i[:, None] * torch.ones(1, 4), (i * 10).to(torch.float64)[:, None, None].expand(-1, 2, 3).contiguous()

    res = jd.sharded_rows(n_items, compute)
    assert calls == ([mine] if mine else [])                  # one call, none for an empty shard
    if rank == 0:
        a, b = res
        want = torch.arange(n_items, dtype=torch.float32)
        assert torch.equal(a, want[:, None] * torch.ones(1, 4))
        assert b.dtype == torch.float64 and torch.equal(b, (want * 10).to(torch.float64)[:, None, None].expand(-1, 2, 3))
    else:
        assert res is None
    one = jd.sharded_rows(n_items, lambda idx: torch.tensor(idx, dtype=torch.int64)[:, None])
    if rank == 0:
        assert torch.equal(one, torch.arange(n_items)[:, None])
    # the bispectrum sweep over ranks, with the GPU call replaced by a stand-in of the same shape contract
    import jax_powspec_b200.correlations as jc

    def fake_pairs(delta, box_size, a, b, theta, **kw):
        a, b, theta = (np.asarray(v, dtype=np.float32) for v in (a, b, theta))
        k_all = np.concatenate([a[:, None], b[:, None], a[:, None] + b[:, None] * np.cos(theta)[None, :]], axis=1)
        return k_all, 2 * k_all, theta, a[:, None] * theta[None, :], b[:, None] * theta[None, :]

    real_pairs, jc.bispec_pairs = jc.bispec_pairs, fake_pairs
    try:
        k1 = np.arange(1, n_items + 1, dtype=np.float32) * 0.01
        k2 = k1 * 2
        theta = np.linspace(0, 3, 4).astype(np.float32)
        got = jd.bispec_pairs_sharded(np.zeros((4, 4, 4), np.float32), 100.0, k1, k2, theta)
        if rank == 0:
            want = fake_pairs(None, 100.0, k1, k2, theta)
            assert isinstance(got[0], np.ndarray)
            for g_, w_ in zip(got, want):
                np.testing.assert_array_equal(g_, w_)
        else:
            assert got is None
    finally:
        jc.bispec_pairs = real_pairs
    res = jd.covariance_batch(list(range(100, 100 + n_items)), _fake_measure)
    if rank == 0:
        rows, mean, cov = res
        np.save(os.path.join(out_dir, "rows.npy"), rows)
        np.save(os.path.join(out_dir, "cov.npy"), cov)
    else:
        assert res is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [7, 1])
def test_covariance_batch_world2(tmp_path, n_items):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, n_items, str(tmp_path)), nprocs=2, join=True)
    rows = np.load(tmp_path / "rows.npy")
    want = np.stack([_fake_measure(s).numpy() for s in range(100, 100 + n_items)])
    np.testing.assert_array_equal(rows, want)
    cov = np.load(tmp_path / "cov.npy")
    flat = want.reshape(n_items, -1).astype(np.float64)
    d = flat - flat.mean(0)
    np.testing.assert_allclose(cov, d.T @ d / max(n_items - 1, 1), rtol=1e-12, atol=1e-15)


def test_single_process_paths():
    from jax_powspec_b200 import dist as jd
    assert jd.world() == (0, 1)
    assert jd.shard_indices(5) == [0, 1, 2, 3, 4]
    assert jd.max_over_ranks(3.5) == 3.5
    res = jd.covariance_batch([1, 2, 3], _fake_measure)
    assert res[0].shape == (3, 5, 3) and res[2].shape == (15, 15)
    a, b = jd.sharded_rows(4, lambda idx: (torch.tensor(idx)[:, None], torch.tensor(idx) * 2))
    assert a.tolist() == [[0], [1], [2], [3]] and b.tolist() == [0, 2, 4, 6]


# --------------------------------------------------------------------------- slab choreography
def _slab_worker(rank, world_size, port, n, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world_size), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    # import only the exchange helpers: the CUDA library is loaded by the package import, which is fine on CPU
    from jax_powspec_b200 import slab
    nxl = n // world_size
    g = torch.Generator().manual_seed(7)
    field = torch.rand((n, n, n), generator=g, dtype=torch.float32)          # the same global field on every rank
    # a "painted" slab whose ghost planes carry contributions that belong to the neighbours
    ghost_contrib = torch.rand((n, n, n), generator=g, dtype=torch.float32)
    lo_plane = torch.rand((n, n, n), generator=g, dtype=torch.float32)
    mesh = torch.zeros((slab.GHOST_LO + nxl + slab.GHOST_HI, n, n))
    x0 = rank * nxl
    mesh[slab.GHOST_LO: slab.GHOST_LO + nxl] = field[x0: x0 + nxl]
    for j in range(slab.GHOST_HI):                                           # what I deposited beyond my slab
        mesh[slab.GHOST_LO + nxl + j] = ghost_contrib[(x0 + nxl + j) % n] * (rank + 1)
    mesh[0] = lo_plane[(x0 - 1) % n] * (rank + 1)
    slab.halo_exchange_add(mesh, nxl)
    # expected owned planes after the exchange
    want = field[x0: x0 + nxl].clone()
    prev, nxt = (rank - 1) % world_size, (rank + 1) % world_size
    for j in range(slab.GHOST_HI):
        want[j] += ghost_contrib[(x0 + j) % n] * (prev + 1)
    want[nxl - 1] += lo_plane[(x0 + nxl - 1) % n] * (nxt + 1)
    owned = mesh[slab.GHOST_LO: slab.GHOST_LO + nxl]
    assert torch.allclose(owned, want, rtol=0, atol=1e-6), "halo exchange landed on the wrong planes"
    # distributed R2C on the ORIGINAL field: rfft2 of owned planes, pack, all-to-all, fft along x
    yz = torch.fft.rfft2(field[x0: x0 + nxl].to(torch.float64)).to(torch.complex64)
    send = slab.pack_blocks_torch(yz, world_size)
    recv = torch.empty_like(send)
    slab.transpose_all_to_all(send, recv)
    dk_local = torch.fft.fft(recv.reshape(n, nxl, n // 2 + 1).to(torch.complex128), dim=0)
    full = torch.fft.rfftn(field.to(torch.float64))
    y0 = rank * nxl
    err = (dk_local - full[:, y0: y0 + nxl, :]).abs().max().item() / full.abs().max().item()
    assert err < 1e-6, f"slab FFT choreography wrong: {err}"
    # particle routing
    gp = torch.Generator().manual_seed(100 + rank)
    xs = torch.rand(1000, generator=gp) * 10.0
    ys = torch.rand(1000, generator=gp)
    rx, ry, rz, rw = slab.route_particles(xs, ys, ys.clone(), None, 10.0, n)
    cell = torch.floor(rx * (n / 10.0)).long() % n
    assert bool(((cell // nxl) == rank).all()) and rw is None and rx.numel() == ry.numel()
    tot = torch.tensor([rx.numel()]); dist.all_reduce(tot)
    assert int(tot.item()) == 1000 * world_size
    sy = torch.tensor([float(ry.sum())], dtype=torch.float64); dist.all_reduce(sy)
    sy0 = torch.tensor([float(ys.sum())], dtype=torch.float64); dist.all_reduce(sy0)
    assert abs(sy.item() - sy0.item()) < 1e-3
    if rank == 0:
        open(os.path.join(out_dir, "ok"), "w").write("ok")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world_size", [2, 4])
def test_slab_choreography_gloo(tmp_path, world_size):
    port = _free_port()
    mp.spawn(_slab_worker, args=(world_size, port, 16, str(tmp_path)), nprocs=world_size, join=True)
    assert (tmp_path / "ok").exists()


def test_bind_near_gpu_is_best_effort_without_nvml():
    """No GPU / no NVML here: the helper must report that and leave the affinity alone."""
    import os
    from jax_powspec_b200.dist import bind_near_gpu
    before = os.sched_getaffinity(0)
    info = bind_near_gpu(0)
    assert info["bound"] is False and "error" in info
    assert os.sched_getaffinity(0) == before
