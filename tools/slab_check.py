"""Real multi-process check + timing of the slab-sharded path (run under torchrun on N GPUs):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29512 \
        tools/slab_check.py [--n-mesh 512] [--n-part 2e7] [--order 4] [--steps 5]
Each rank generates uniform particles inside its own x-slab (BASELINE configs[3]); rank 0 also runs
the single-GPU pipeline on the gathered catalogue when it fits (--check) and compares."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument("--n-mesh", type=int, default=512)
ap.add_argument("--n-part", type=float, default=2e7)
ap.add_argument("--order", type=int, default=4)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--box", type=float, default=2000.0)
ap.add_argument("--check", action="store_true")
ap.add_argument("--method", default="auto")
ap.add_argument("--transport", default="auto")
ap.add_argument("--no-overlap", action="store_true")
ap.add_argument("--layout", default="auto")
a = ap.parse_args()

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
import jax_powspec_b200 as jps
from jax_powspec_b200 import _lib
from jax_powspec_b200.slab import SlabPipeline

n, box = a.n_mesh, a.box
npart = int(a.n_part); nloc = npart // world
g = torch.Generator(device=dev); g.manual_seed(42 + rank)
w_slab = box / world
x = (torch.rand(nloc, generator=g, device=dev) * w_slab + rank * w_slab).clamp_(max=np.nextafter(np.float32((rank + 1) * w_slab), np.float32(0)))
y = torch.rand(nloc, generator=g, device=dev) * box; y[y >= box] = 0
z = torch.rand(nloc, generator=g, device=dev) * box; z[z >= box] = 0
kF = 2 * np.pi / box
ke = np.arange(kF, np.pi * n / box, kF).astype(np.float32)
pipe = SlabPipeline(n, box, ke, order=a.order, compat="fixed", method=a.method, transport=a.transport,
                    overlap=not a.no_overlap, layout=a.layout)

def sync():
    if world > 1: dist.barrier()
    torch.cuda.synchronize()

for _ in range(3): pipe(x, y, z)
sync()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps): k3d, pk, nm = pipe(x, y, z)
e1.record(); sync()
ms = e0.elapsed_time(e1) / a.steps
t = torch.tensor([ms], dtype=torch.float64, device=dev)
if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
ms = float(t.item())
# per-stage times (events around each stage, one extra pass)
stages = {}
def timed(name, fn):
    sync(); s0 = torch.cuda.Event(enable_timing=True); s1 = torch.cuda.Event(enable_timing=True)
    s0.record(); fn(); s1.record(); sync(); stages[name] = s0.elapsed_time(s1)
from jax_powspec_b200.slab import halo_exchange_add, transpose_all_to_all
timed("paint", lambda: pipe.stage_paint(x, y, z))
if world > 1: timed("halo", lambda: halo_exchange_add(pipe.mesh, pipe.nxl))
timed("fft_yz_pack", pipe.stage_fft_yz_pack)
if world > 1 and pipe.transport == "nccl": timed("all_to_all", lambda: transpose_all_to_all(pipe.buf_b, pipe.buf_a))
timed("fft_x", pipe.stage_fft_x)
timed("bin", lambda: pipe.stage_partial(True))
res = {"transport": pipe.transport, "xfast": pipe.xfast, "p2p_error": getattr(pipe, "_p2p_error", None), "n_gpus": world, "n_mesh": n, "n_part": npart, "order": a.order, "ms_per_step": ms,
       "gparticles_per_s": npart / ms / 1e6, "stages_ms_rank0": stages,
       "a2a_bytes_per_rank": (world - 1) / world ** 2 * 8 * n * n * (n // 2 + 1)}
if a.check:
    pk_d = pk.cpu().numpy(); nm_d = nm.cpu().numpy()
    parts = [torch.empty(nloc, device=dev) for _ in range(world)] if world > 1 else None
    full = []
    for tns in (x, y, z):
        if world > 1:
            dist.all_gather(parts, tns); full.append(torch.cat(parts))
        else:
            full.append(tns)
    if rank == 0:
        k1, pk1, nm1 = jps.paint_powspec(full[0], full[1], full[2], None, 0., 0., 0., box, n, ke, order=a.order, compat="fixed")
        pk1 = pk1.cpu().numpy(); nm1 = nm1.cpu().numpy()
        err = np.abs(pk_d - pk1) / np.abs(pk1[:, :1])
        res["check"] = {"counts_equal": bool(np.array_equal(nm_d, nm1)), "max_rel_to_P0": float(np.nanmax(err))}
if rank == 0:
    print(json.dumps(res), flush=True)
if world > 1: dist.destroy_process_group()
