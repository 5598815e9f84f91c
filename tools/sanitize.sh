#!/bin/bash
# compute-sanitizer passes over the hand-written kernels at small sizes (race detection).
# Run on the GPU box:   gpurun --timeout 900 -- 'bash tools/sanitize.sh'
# memcheck: out-of-bounds / misaligned accesses; racecheck: shared-memory hazards in the tile painters,
# the bucketing passes, the binning kernels and the block scans of the mock generator.
# Each pass runs a SMALL selection of the GPU tests (the sanitizer slows kernels down 10-100x).
set -u
mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
SEL='test_known_answers or test_translation_by_whole_cells or test_plane_wave_known_answer or test_golden_bispec or test_populate_device_in_device_out_and_empty or test_golden_gaussian_field or test_interlacing_leaves_a_band_limited_field_alone'
for tool in memcheck racecheck; do
  timeout 600 $SAN --tool $tool --error-exitcode 99 --print-limit 20 python tools/sanitize_paint.py \
      > gpurun_out/sanitize_paint_$tool.log 2>&1
  echo "$tool exit code $?" >> gpurun_out/sanitize_paint_$tool.log
  tail -6 gpurun_out/sanitize_paint_$tool.log
done
# both shapes of the staged fine pass (128 tiles per group at 256^3), the bank-class order of the deposit, and the tile
# histogram fused into the coarse pass (N = 592)
for tool in memcheck racecheck; do
  for cfg in "JPS_FINE_CHUNK=small JPS_COUNT=separate" "JPS_FINE_CHUNK=big JPS_TILE_ORDER=bank"; do
    env JPS_BUCKET=two JPS_FINE=staged JPS_MAX_GROUPS=64 $cfg timeout 600 $SAN --tool $tool --error-exitcode 99 --print-limit 20 \
        python tests/helpers/two_level_check.py --quick >> gpurun_out/sanitize_fine_$tool.log 2>&1
    echo "[$cfg] $tool exit code $?" >> gpurun_out/sanitize_fine_$tool.log
  done
  env JPS_COUNT=fused JPS_FINE=staged timeout 600 $SAN --tool $tool --error-exitcode 99 --print-limit 20 \
      python tests/helpers/big_mesh_check.py --quick >> gpurun_out/sanitize_fine_$tool.log 2>&1
  echo "[big mesh, fused count] $tool exit code $?" >> gpurun_out/sanitize_fine_$tool.log
  tail -10 gpurun_out/sanitize_fine_$tool.log
done
for tool in memcheck racecheck; do
  timeout 400 $SAN --tool $tool --error-exitcode 99 --print-limit 20 \
      python -m pytest tests -m gpu -q -x -p no:cacheprovider -k "$SEL" \
      > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool exit code $?" >> gpurun_out/sanitize_$tool.log
  tail -5 gpurun_out/sanitize_$tool.log
done
