#!/bin/bash
# A/B of the painter on ONE rank's share of the 8-GPU C4 decomposition (1.25e8 particles, 256 planes of 2048^3, PCS)
# and on C2: round-1 paths (JPS_TILE_FLUSH=red, JPS_FINE_TWOPASS=1) against the current ones.
mkdir -p gpurun_out
for cfg in "JPS_TILE_FLUSH=red JPS_FINE_TWOPASS=1" "JPS_TILE_FLUSH=red" "JPS_FINE_TWOPASS=1" ""; do
  echo "== C4 rank [$cfg]"; env $cfg python tools/slab_rank_paint_profile.py
done 2>&1 | tee gpurun_out/r2_rank_paint_ab.log
for cfg in "JPS_TILE_FLUSH=red" ""; do
  echo "== C2 [$cfg]"; env $cfg python bench.py --workload c2 --quick-kernels --steps 5
done 2>&1 | tee -a gpurun_out/r2_rank_paint_ab.log
