"""The slab-sharded path over REAL processes: torch.distributed.run with one rank per GPU, NCCL for the
exchanges and CUDA-IPC peer memory for the fused transposing peer-store.  Runs at every world size in
{2, 4, 8} the box offers and is skipped on a one-GPU box (there tests/test_gpu_slab.py drives the same
kernels with virtual ranks).  Mode counts must equal the single-GPU pipeline's bit for bit and the
multipoles agree to 1e-5 of P0 in every (transport, layout, overlap) combination."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 4, 8])
def test_slab_pipeline_over_real_ranks(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, box has {torch.cuda.device_count()}")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "helpers", "multi_rank_check.py"), "256", "4e6"]
    env = dict(os.environ)
    if world == 2:
        env["JPS_SLAB_CHUNKS"] = "4"      # 128 owned planes -> 4 pieces of 32: the TMA bulk-store peer kernel takes them
    r = subprocess.run(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-4000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("MULTI_RANK_RESULT ")]
    assert line, r.stdout[-4000:]
    res = json.loads(line[-1][len("MULTI_RANK_RESULT "):])
    assert res["world"] == world and len(res["cases"]) == 15
    for c in res["cases"]:
        assert c["counts_equal"] and c["k_equal"], c
        assert c["max_rel_P0"] <= 1e-5, c
        if "host_pipeline_max_rel" in c:
            assert c["host_pipeline_max_rel"] <= 1e-5, c
    assert {c["transport"] for c in res["cases"]} == {"p2p", "nccl"}, "the peer-memory transport must have been exercised"
    assert any(c["pipelined"] for c in res["cases"]), "the deposit / FFT / transfer pipeline must have been exercised"
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, f"multi_rank_parity_w{world}.json"), "w") as f:
        json.dump(res, f, indent=1)
