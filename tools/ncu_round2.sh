#!/bin/bash
# ncu evidence of round 2 (run on a ONE-GPU box; ncu replays every profiled kernel ~40 times):
#  1. launch list of the default bench (C4 on one GPU), timed region only
#  2. --set full capture of one launch of each painter kernel on ONE rank's share of the 8-GPU C4 decomposition
#  3. the same for C2 (deposit + binning)
mkdir -p gpurun_out
NCU=/usr/local/cuda/bin/ncu
$NCU --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "jps_timed/" -c 400 --csv \
     --log-file gpurun_out/r2_launches_c4_1gpu.csv python bench.py --quick --steps 2 --warmup 3 > gpurun_out/r2_ncu_c4_1gpu.log 2>&1
tail -2 gpurun_out/r2_ncu_c4_1gpu.log
# 3 warm-up paints x 4 matching kernels are skipped
$NCU --set full --clock-control none --import-source on -k regex:'paint_tile_fx|coarse_scatter|fine_scatter|bucket_count' \
     --launch-skip 12 --launch-count 4 -f -o gpurun_out/r2_prof_c4_rank python tools/slab_rank_paint_profile.py > gpurun_out/r2_ncu_c4_rank.log 2>&1
tail -2 gpurun_out/r2_ncu_c4_rank.log
$NCU --set full --clock-control none --import-source on -k regex:'paint_tile_fx|pk_fold_bin|fine_scatter|coarse_scatter' \
     --nvtx --nvtx-include "jps_timed/" --launch-count 4 -f -o gpurun_out/r2_prof_c2 python bench.py --workload c2 --quick --steps 1 --warmup 3 > gpurun_out/r2_ncu_c2.log 2>&1
tail -2 gpurun_out/r2_ncu_c2.log
ls -la gpurun_out/*.ncu-rep
