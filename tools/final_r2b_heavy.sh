#!/bin/bash
# Validation of the heavy-group slices of the fine pass (ONE GPU): full GPU test suite, the clustered catalogues, sanitizer
# on the heavy tests, and the timings that must not move.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x --durations=4 ) > gpurun_out/r2b_tests_heavy.log 2>&1
tail -8 gpurun_out/r2b_tests_heavy.log
timeout 300 python tools/clustered_paint.py 2>&1 | tee gpurun_out/r2b_clustered_paint_heavy.log | cut -c1-420
SAN=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck; do
  timeout 300 $SAN --tool $tool --error-exitcode 99 --print-limit 20 python -m pytest tests/test_gpu_paint.py -m gpu -q -x -p no:cacheprovider \
      -k "heavy_group" > gpurun_out/r2b_sanitize_heavy_$tool.log 2>&1
  echo "$tool exit code $?" >> gpurun_out/r2b_sanitize_heavy_$tool.log; tail -3 gpurun_out/r2b_sanitize_heavy_$tool.log
done
timeout 300 python tools/slab_rank_paint_profile.py 2>&1 | tail -1 | tee gpurun_out/r2b_rank8_heavy.log
timeout 300 python tools/paint_profile.py --reps 2 2>&1 | tail -1 | tee gpurun_out/r2b_paint_1gpu_heavy.log
