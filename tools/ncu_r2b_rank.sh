#!/bin/bash
# ncu --set full of the four painter kernels on one rank's share of the 8-GPU C4 decomposition (final round-2b tree);
# refreshes profiles/r2_ncu_full_summary_c4_rank.md and r2_ncu_dram_traffic_c4_rank.json (copies land in gpurun_out/).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NCU=/usr/local/cuda/bin/ncu
timeout 600 $NCU --set full --clock-control none --import-source on -k regex:'paint_tile_fx|coarse_scatter|fine_staged|fine_scatter|bucket_count' \
     --launch-skip 12 --launch-count 4 -f -o gpurun_out/r2b_prof_c4_rank python tools/slab_rank_paint_profile.py > gpurun_out/r2b_ncu_c4_rank.log 2>&1
tail -2 gpurun_out/r2b_ncu_c4_rank.log
$NCU -i gpurun_out/r2b_prof_c4_rank.ncu-rep --page raw --csv > gpurun_out/r2b_raw_c4_rank.csv 2>/dev/null
python tools/make_profiles.py --raw gpurun_out/r2b_raw_c4_rank.csv --tag r2 --suffix c4_rank \
     --title "one rank of the 8-GPU C4 decomposition (1.25e8 uniform particles, PCS, 259 planes of 2048^2; tools/slab_rank_paint_profile.py), final round-2b tree"
cp profiles/r2_ncu_full_summary_c4_rank.md profiles/r2_ncu_dram_traffic_c4_rank.json gpurun_out/
