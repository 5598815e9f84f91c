#!/bin/bash
# Last GPU call of round 1: the whole GPU suite on the final tree, then the 1-GPU point of the C4 series.
mkdir -p gpurun_out
( timeout 110 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -25 ) > gpurun_out/final_gpu_tests.log 2>&1
( timeout 70 python bench.py --workload c4 --gpus 1 --steps 3 --warmup 3 2>&1 | tail -5 ) > gpurun_out/c4_1gpu.log 2>&1
