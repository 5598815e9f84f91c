"""Multi-process parity check of the slab-sharded path over REAL NCCL / CUDA-IPC peer memory (run
under torch.distributed.run by tests/test_gpu_multi.py; one process per GPU).

Every rank starts from a random 1/W share of ONE seeded clustered catalogue (not slab-sorted), routes
its particles to the slab owners (variable-size all-to-all), and runs SlabPipeline in every
(transport, layout, overlap) combination; rank 0 also runs the single-GPU pipeline on the whole
catalogue.  Prints one JSON line: per combination counts_equal and max |dP|/P0."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
import torch.distributed as dist

from tests.util import clustered_particles

world = int(os.environ["WORLD_SIZE"]); rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
import jax_powspec_b200 as jps
from jax_powspec_b200.slab import SlabHostPipeline, SlabPipeline, route_particles

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
npart = int(float(sys.argv[2])) if len(sys.argv) > 2 else 2_000_000
box = 1000.0
p = clustered_particles(1234, npart, box)                      # identical on every rank (seeded NumPy)
w_all = (np.random.default_rng(5).random(npart).astype(np.float32) + np.float32(0.5))
kF = 2 * np.pi / box
ke = np.arange(kF, np.pi * n / box, kF).astype(np.float32)
perm = np.random.default_rng(99).permutation(npart)
mine = perm[rank::world]                                       # a random share, NOT slab-sorted
res = {"world": world, "n_mesh": n, "n_part": npart, "cases": []}
for order, weighted in ((4, False), (3, True), (2, False)):
    xs, ys, zs = (torch.from_numpy(np.ascontiguousarray(p[mine, i])).to(dev) for i in range(3))
    ws = torch.from_numpy(w_all[mine]).to(dev) if weighted else None
    x, y, z, w = route_particles(xs, ys, zs, ws, box, n)
    total = torch.tensor([x.numel()], device=dev)
    dist.all_reduce(total)
    assert int(total.item()) == npart, "route_particles lost particles"
    if rank == 0:
        full = [torch.from_numpy(np.ascontiguousarray(p[:, i])).to(dev) for i in range(3)]
        wf = torch.from_numpy(w_all).to(dev) if weighted else None
        ref = jps.PaintPowspec(n, box, ke, order=order, compat="fixed", device=dev)
        k1, pk1, nm1 = (t.clone() for t in ref(*full, wf))
    for transport, layout, overlap, pipelined, fft in (("p2p", "xfast", True, True, "pencil"), ("p2p", "xfast", True, False, "cufft2d"),
                                                       ("p2p", "xfast", False, False, "pencil"), ("p2p", "xslow", True, True, "auto"),
                                                       ("nccl", "xslow", True, False, "auto")):
        pipe = SlabPipeline(n, box, ke, order=order, compat="fixed", transport=transport, layout=layout, overlap=overlap,
                            pipeline=pipelined, fft=fft)
        pipe._force_chunks = overlap
        for _ in range(2):                                      # twice: buffers reused across steps
            k3d, pk, nm = pipe(x, y, z, w)
        case = {"order": order, "weighted": weighted, "transport": pipe.transport, "layout": "xfast" if pipe.xfast else "xslow",
                "overlap": overlap, "fft": fft, "pipelined": bool(pipelined and pipe._can_pipeline(x.numel()))}
        if transport == "p2p" and layout == "xfast" and overlap and pipelined and order == 4:
            # the host-buffer pipeline (what bench.py's e2e times) must give the same numbers
            host = SlabHostPipeline(pipe, x.numel(), weighted=weighted, n_chunks=3)
            hk, hpk, hnm = host(x.cpu().numpy(), y.cpu().numpy(), z.cpu().numpy(), None if w is None else w.cpu().numpy())
            case["host_pipeline_max_rel"] = float(np.nanmax(np.abs(hpk - pk.cpu().numpy()) / np.abs(pk.cpu().numpy()[:, :1])))
        if rank == 0:
            err = ((pk - pk1).abs() / pk1[:, :1].abs()).max().item()
            case.update(counts_equal=bool(torch.equal(nm, nm1)), k_equal=bool(torch.equal(k3d, k1)), max_rel_P0=float(err))
        res["cases"].append(case)
        pipe.close()
        dist.barrier()
if rank == 0:
    print("MULTI_RANK_RESULT " + json.dumps(res), flush=True)
dist.destroy_process_group()
