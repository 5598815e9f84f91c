import os, sys, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
world = int(os.environ["WORLD_SIZE"]); rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
import jax_powspec_b200 as jps
from jax_powspec_b200.slab import SlabPipeline
def log(*a): print(f"[r{rank}]", *a, flush=True)
n, box = 64, 1000.0
ke = np.arange(0.01, 0.19, 0.01).astype(np.float32)
pipe = SlabPipeline(n, box, ke, order=2, transport="auto")
log("transport", pipe.transport, getattr(pipe, "_p2p_error", None))
if pipe.transport == "p2p":
    log("peer ptrs", [hex(int(p or 0)) for p in pipe.peer_ptrs])
    try:
        pipe.buf_b.fill_(complex(10 + rank, 0))
        from jax_powspec_b200._lib import lib, check
        from jax_powspec_b200.plan import ptr, stream_ptr
        check(lib.jps_slab_pack_p2p(pipe.handle, ptr(pipe.buf_b), pipe.peer_ptrs, stream_ptr()), "pack_p2p")
        torch.cuda.synchronize(); dist.barrier()
        log("after pack_p2p: slots", [pipe.buf_a[q, 0, 0, 0].item() for q in range(world)])
    except Exception:
        traceback.print_exc()
dist.barrier(); dist.destroy_process_group()
