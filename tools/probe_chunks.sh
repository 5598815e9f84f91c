#!/bin/bash
# pieces of the FFT / peer-transfer overlap (JPS_SLAB_CHUNKS) on N GPUs: stage times of the default bench
N=${1:-2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for c in 8 16 32; do
  JPS_SLAB_CHUNKS=$c timeout 300 python bench.py --gpus $N --quick-kernels --steps 5 --no-cpu 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); t=d.get('transpose') or {}
print('chunks=$c', 'step', round(d['ms_per_step'],2), 'stages', {k:round(v,2) for k,v in (d.get('stages_ms') or d.get('stages_ms_max_over_ranks') or {}).items()}, 'fft_alone', t.get('fft_yz_alone_ms'), 'store_alone', t.get('store_alone_ms'), 'hidden', t.get('hidden_ms'))" | tee -a gpurun_out/r2b_chunks_${N}gpu.log
done
