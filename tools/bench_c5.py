"""BASELINE.json configs[4] ("C5") as stated: P(k) covariance batch of 256 lognormal realisations of ~1e7
particles, CIC on 512^3, k_edges = arange(0.003, k_Ny, 0.0025) (/root/reference/tests/voids.py:55), throughput in
realisations/s.  Realisations are independent units: seed s goes to rank s % W (no data-path collective), and on
every GPU `--streams` host threads drive their OWN PaintPowspec pipeline (own plan, own bucketing workspace) on
their own CUDA stream, so the mock generation of one realisation (device generator, one 8-byte host read-back)
overlaps the painting of another.  Rank 0 gathers the 256 x (nbins x 3) rows and forms the sample covariance.

    python tools/bench_c5.py [--seeds 256] [--streams 2]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tools/bench_c5.py
Prints one JSON line on rank 0."""
import argparse, json, os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument("--seeds", type=int, default=256)
ap.add_argument("--streams", type=int, default=2)
ap.add_argument("--n-mesh", type=int, default=512)
ap.add_argument("--n-part", type=float, default=1e7)
a = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
import jax_powspec_b200 as jps
from jax_powspec_b200 import dist as jd, mocks

n, box, npart = a.n_mesh, 1000.0, int(a.n_part)
ke = np.arange(0.003, np.pi * n / box, 0.0025).astype(np.float32)
kf_t = np.linspace(1e-4, 10, 4056)
pk_t = 2.0e4 * (kf_t / 0.02) / (1.0 + (kf_t / 0.02) ** 2) ** 1.7
dens = npart / box ** 3
mine = jd.shard_indices(a.seeds)
rows = {}
nparts = {}


def worker(tid):
    torch.cuda.set_device(local)
    stream = torch.cuda.Stream(dev)
    with torch.cuda.stream(stream):
        pipe = jps.PaintPowspec(n, box, ke, order=2, compat="fixed", n_part_max=int(npart * 1.1), device=dev)
        for seed in mine[tid::a.streams]:
            p = mocks.lognormal_mock(256, kf_t, pk_t, 1.1, dens, seed, box)        # device generator (row f-3)
            nparts[seed] = int(p.shape[0])
            rows[seed] = pipe(p[:, 0], p[:, 1], p[:, 2])[1].clone()
        stream.synchronize()


def run():
    th = [threading.Thread(target=worker, args=(t,)) for t in range(a.streams)]
    for t in th: t.start()
    for t in th: t.join()

# warm-up: one realisation per stream (plans, bin tables, cuFFT kernels)
save, mine = mine, mine[: a.streams]
run()
mine = save
rows.clear(); nparts.clear()
if world > 1: dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
run()
torch.cuda.synchronize()
if world > 1: dist.barrier()
dt = time.perf_counter() - t0
local_rows = torch.stack([rows[s].reshape(-1) for s in mine]) if mine else torch.zeros((0, (len(ke) - 1) * 3), device=dev)
full = jd.gather_rows(local_rows, a.seeds)
t = torch.tensor([dt], dtype=torch.float64, device=dev)
if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    arr = full.cpu().numpy().astype(np.float64)
    ok = np.isfinite(arr).all(axis=0)
    mean, cov = jd.sample_covariance(arr[:, ok])
    print(json.dumps({"config": "C5: 256 lognormal realisations, CIC 512^3, P0/P2/P4", "n_gpus": world, "streams_per_gpu": a.streams,
                      "realisations": a.seeds, "particles_mean": float(np.mean(list(nparts.values()))), "seconds": float(t.item()),
                      "realisations_per_s": a.seeds / float(t.item()), "includes": "mock generation on the device + paint + FFT + multipoles",
                      "cov_shape": list(cov.shape), "cov_finite": bool(np.isfinite(cov).all()),
                      "P0_mean_first_bins": [float(v) for v in arr[:, :9:3].mean(axis=0)]}), flush=True)
if world > 1: dist.destroy_process_group()
