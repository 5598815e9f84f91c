#!/bin/bash
# Round-2b probe 4 (ONE GPU): ATOMS throughput against the bank pattern, and the deposit's shared-memory wavefronts with
# the particles of a tile in bank-class order / in arrival order.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
./tools/microbench_atoms | tee gpurun_out/r2b_microbench_atoms.txt
NCU=/usr/local/cuda/bin/ncu
M=l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__inst_executed.sum,gpu__time_duration.sum,smsp__inst_executed_op_shared_atom.sum,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active
for cfg in "JPS_TILE_ORDER=bank" "JPS_TILE_ORDER=arrival"; do
  echo "== rank8 deposit [$cfg]"
  env $cfg timeout 300 $NCU --metrics $M --clock-control none -k regex:paint_tile_fx --launch-skip 3 --launch-count 1 \
      python tools/slab_rank_paint_profile.py 2>&1 | grep -E "paint_tile_fx|l1tex|smsp|gpu__time" 
done | tee gpurun_out/r2b_deposit_order_ncu.txt
