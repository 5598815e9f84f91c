#!/bin/bash
# Validation of the heavy-tile split (ONE GPU): full GPU test suite, compute-sanitizer on the heavy-tile test, the clustered
# catalogues, and the painter / bench timings that must not move.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x --durations=4 ) > gpurun_out/r2b_tests_split.log 2>&1
tail -9 gpurun_out/r2b_tests_split.log
SAN=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck; do
  timeout 400 $SAN --tool $tool --error-exitcode 99 --print-limit 20 python -m pytest tests/test_gpu_paint.py -m gpu -q -x -p no:cacheprovider \
      -k "heavy_tiles" > gpurun_out/r2b_sanitize_split_$tool.log 2>&1
  echo "$tool exit code $?" >> gpurun_out/r2b_sanitize_split_$tool.log; tail -3 gpurun_out/r2b_sanitize_split_$tool.log
done
timeout 300 python tools/clustered_paint.py 2>&1 | tee gpurun_out/r2b_clustered_paint_split.log | cut -c1-420
timeout 300 python tools/slab_rank_paint_profile.py 2>&1 | tail -1 | tee gpurun_out/r2b_rank8_split.log
timeout 300 python bench.py --workload c2 --quick-kernels --steps 5 2>&1 | tail -1 | cut -c1-900 | tee gpurun_out/r2b_c2_split.log
timeout 600 python bench.py > gpurun_out/r2b_bench_c4_1gpu_split.json 2> gpurun_out/r2b_bench_c4_1gpu_split.err; tail -1 gpurun_out/r2b_bench_c4_1gpu_split.json | cut -c1-300
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
