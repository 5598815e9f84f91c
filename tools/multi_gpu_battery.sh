#!/bin/bash
# Everything that needs N GPUs, in one gpurun call:  gpurun --gpus N -- 'bash tools/multi_gpu_battery.sh N'
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/r2_topo_${N}gpu.txt 2>&1
(time python -m pytest tests/test_gpu_multi.py -m gpu -x -q) > gpurun_out/r2_tests_multi_${N}gpu.log 2>&1; tail -6 gpurun_out/r2_tests_multi_${N}gpu.log
python bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_c4_${N}gpu.json 2> gpurun_out/r2_bench_c4_${N}gpu.err
tail -c 600 gpurun_out/r2_bench_c4_${N}gpu.err; cut -c1-400 gpurun_out/r2_bench_c4_${N}gpu.json
python bench.py --gpus $N --quick-kernels --steps 5 --pipeline > gpurun_out/r2_c4_${N}gpu_pipeline.json 2>&1; cut -c1-600 gpurun_out/r2_c4_${N}gpu_pipeline.json
JPS_PACK_KERNEL=tma python bench.py --gpus $N --quick-kernels --steps 5 > gpurun_out/r2_c4_${N}gpu_tma_pack.json 2>&1; cut -c1-600 gpurun_out/r2_c4_${N}gpu_tma_pack.json
JPS_SLAB_CHUNKS=4 python bench.py --gpus $N --quick-kernels --steps 5 > gpurun_out/r2_c4_${N}gpu_4chunks.json 2>&1; cut -c1-300 gpurun_out/r2_c4_${N}gpu_4chunks.json
$TR --master-port 29533 tools/bench_c5.py > gpurun_out/r2_c5_${N}gpu.json 2>&1; tail -1 gpurun_out/r2_c5_${N}gpu.json
$TR --master-port 29544 tools/bench_bispec_sharded.py > gpurun_out/r2_bispec_sharded_${N}gpu.json 2>&1; tail -1 gpurun_out/r2_bispec_sharded_${N}gpu.json
