#!/bin/bash
# Everything that needs N GPUs, in one gpurun call:  gpurun --gpus N -- 'bash tools/multi_gpu_battery.sh N'
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/r2_topo_${N}gpu.txt 2>&1
(time python -m pytest tests/test_gpu_multi.py -m gpu -x -q) > gpurun_out/r2_tests_multi_${N}gpu.log 2>&1; tail -6 gpurun_out/r2_tests_multi_${N}gpu.log
python bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_c4_${N}gpu.json 2> gpurun_out/r2_bench_c4_${N}gpu.err
tail -c 600 gpurun_out/r2_bench_c4_${N}gpu.err; cut -c1-400 gpurun_out/r2_bench_c4_${N}gpu.json
summ() { python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{\"ms_per_step'):
        d=json.loads(l); t=d['transpose'] or {}
        print('$1', 'step', round(d['ms_per_step'],2), 'stages', {k:round(v,2) for k,v in d['stages_ms'].items()}, 'fft_alone', t.get('fft_yz_alone_ms'), 'store_alone', t.get('store_alone_ms'), 'hidden', t.get('hidden_ms'))"; }
python bench.py --gpus $N --quick-kernels --steps 5 --pipeline 2>/dev/null | tee gpurun_out/r2_c4_${N}gpu_pipeline.json | summ pipeline
for v in 2 6; do JPS_PACK_CTAS_PER_SM=$v python bench.py --gpus $N --quick-kernels --steps 5 2>/dev/null | tee gpurun_out/r2_c4_${N}gpu_ctas$v.json | summ ctas_per_sm=$v; done
JPS_SLAB_FFT=cufft2d python bench.py --gpus $N --quick-kernels --steps 5 2>/dev/null | tee gpurun_out/r2_c4_${N}gpu_cufft2d.json | summ cufft2d
$TR --master-port 29533 tools/bench_c5.py > gpurun_out/r2_c5_${N}gpu.json 2>&1; tail -1 gpurun_out/r2_c5_${N}gpu.json
$TR --master-port 29544 tools/bench_bispec_sharded.py > gpurun_out/r2_bispec_sharded_${N}gpu.json 2>&1; tail -1 gpurun_out/r2_bispec_sharded_${N}gpu.json
