"""BASELINE.json configs[2] (all triangle bins of a 256^3 CIC mesh) over several GPUs: every rank holds the mesh
and evaluates its round-robin share of the 276 (k1 <= k2) pairs (dist.bispec_pairs_sharded; no mesh sharding, no
data-path collective, rank 0 gathers the rows).  Prints one JSON line on rank 0.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29544 tools/bench_bispec_sharded.py"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
import jax_powspec_b200 as jps
from jax_powspec_b200 import dist as jd
from jax_powspec_b200.mocks import lognormal_catalog
n, box, npart = 256, 1000.0, 3_500_000
x, y, z = lognormal_catalog(npart, box, n_grid=128, seed=5, device=dev)              # same mesh on every rank
rho = jps.cic_mas_vec(torch.zeros((n, n, n), device=dev), x, y, z, None, npart, 0., 0., 0., box, n, True)
delta = rho / rho.mean() - 1.0
theta = np.linspace(0, np.pi, 20).astype(np.float32)
kF = 2 * np.pi / box
k1s, k2s = jps.triangle_pairs(np.arange(2 * kF, 0.3, 2 * kF).astype(np.float32))
times = []
for it in range(3):                                                                  # first pass is cold (indicator sums)
    if world > 1: dist.barrier()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res = jd.bispec_pairs_sharded(delta, box, k1s, k2s, theta)
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    times.append(time.perf_counter() - t0)
if rank == 0:
    ref = jps.bispec_pairs(delta, box, k1s, k2s, theta)
    B, Bref = res[3], ref[3]
    print(json.dumps({"config": "C3 all triangles: 276 pairs x 20 angles, 256^3", "n_gpus": world, "cold_s": times[0],
                      "warm_s": min(times[1:]), "rows_equal_single_gpu": bool(torch.allclose(B, Bref, rtol=1e-5, atol=1e-6 * float(Bref.abs().max())))}), flush=True)
if world > 1: dist.destroy_process_group()
