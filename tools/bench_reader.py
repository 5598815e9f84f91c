"""C1-sized catalogue text (5e6 rows x 4 columns, '%.6f') -> device float32: GPU reader vs np.loadtxt / pandas."""
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jax_powspec_b200 import _lib, reader  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 5_000_000
rng = np.random.default_rng(0)
p = rng.uniform(0, 2500, (n, 4))
d = tempfile.mkdtemp()
path = os.path.join(d, "cat.dat")
t0 = time.time()
import pandas as pd
pd.DataFrame(p).to_csv(path, sep=" ", header=False, index=False, float_format="%.6f")
nbytes = os.path.getsize(path)
res = {"rows": n, "file_bytes": nbytes, "write_s": time.time() - t0}
reader.read_catalog_text(path, box_size=2500.0)          # warm-up (pinned allocation, module load)
torch.cuda.synchronize()
t0 = time.time()
out, info = reader.read_catalog_text(path, box_size=2500.0, return_info=True)
torch.cuda.synchronize()
res["gpu_reader_total_s"] = time.time() - t0
host, _ = reader._pinned_file(path)
text = host.cuda()
torch.cuda.synchronize()
_lib.profile_enable(True)
_lib.profile_reset()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ev[0].record()
for _ in range(5):
    o = reader.parse_catalog_bytes(text, nbytes, box_size=2500.0)
ev[1].record()
torch.cuda.synchronize()
res["gpu_parse_device_resident_ms"] = ev[0].elapsed_time(ev[1]) / 5
res["gpu_parse_gbs_of_text"] = nbytes / (res["gpu_parse_device_resident_ms"] * 1e-3) / 1e9
res["kernels"] = {k: {"launches": c, "ms": ms} for k, (c, ms) in _lib.profile_snapshot().items()}
m = min(n, 500_000)
small = os.path.join(d, "small.dat")
with open(path, "rb") as f, open(small, "wb") as g:
    for _ in range(m):
        g.write(f.readline())
t0 = time.time()
ref = np.loadtxt(small, usecols=(0, 1, 2), dtype=np.float32)
res["np_loadtxt_s_per_5e6_rows"] = (time.time() - t0) * 5e6 / m
t0 = time.time()
pdref = pd.read_csv(path, usecols=(0, 1, 2), sep=r"\s+", engine="c", header=None).values.astype(np.float32)
res["pandas_read_csv_s"] = time.time() - t0
keep = ((pdref < 2500.0) & (pdref > 0)).all(axis=1)
res["matches_pandas_bits"] = bool(np.array_equal(out.cpu().numpy().view(np.uint32), pdref[keep].view(np.uint32)))
res["matches_loadtxt_bits_first_rows"] = bool(np.array_equal(
    reader.read_catalog_text(small).cpu().numpy().view(np.uint32), ref.view(np.uint32)))
print(json.dumps(res))
