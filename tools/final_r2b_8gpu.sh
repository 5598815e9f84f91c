#!/bin/bash
# Round-2b on N GPUs (default 8): the default bench line and the multi-process parity test.
N=${1:-8}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2b_bench_c4_${N}gpu.json 2> gpurun_out/r2b_bench_c4_${N}gpu.err
tail -c 400 gpurun_out/r2b_bench_c4_${N}gpu.err; tail -1 gpurun_out/r2b_bench_c4_${N}gpu.json | cut -c1-600
(time timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q) > gpurun_out/r2b_tests_multi_${N}gpu.log 2>&1; tail -6 gpurun_out/r2b_tests_multi_${N}gpu.log
