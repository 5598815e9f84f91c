#!/bin/bash
# Final one-GPU validation of round 2b: GPU tests, compute-sanitizer, ncu evidence at C4's size on one GPU, bench lines.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NCU=/usr/local/cuda/bin/ncu
( time timeout 900 python -m pytest tests -m gpu -q -x --durations=6 ) > gpurun_out/r2b_tests_final.log 2>&1
tail -12 gpurun_out/r2b_tests_final.log
timeout 1500 bash tools/sanitize.sh > gpurun_out/r2b_sanitize.log 2>&1
grep -h "exit code\|ERROR SUMMARY\|RACECHECK SUMMARY" gpurun_out/sanitize_*.log | sort | uniq -c
# ncu --set full: the four painter kernels at C4's size on ONE GPU (painter alone: mesh + records + catalogue = 78 GB,
# lean enough for ncu's save / restore); 2 warm-up paints x 4 matching kernels are skipped
timeout 900 $NCU --set full --clock-control none --import-source on -k regex:'paint_tile_fx|coarse_scatter|fine_staged|fine_scatter|bucket_count' \
     --launch-skip 8 --launch-count 4 -f -o gpurun_out/r2b_prof_c4_1gpu python tools/paint_profile.py --warmup 2 --reps 1 > gpurun_out/r2b_ncu_c4_1gpu.log 2>&1
tail -2 gpurun_out/r2b_ncu_c4_1gpu.log
$NCU -i gpurun_out/r2b_prof_c4_1gpu.ncu-rep --page raw --csv > gpurun_out/r2b_raw_c4_1gpu.csv 2>/dev/null
# launch list of the default bench's timed region
timeout 600 $NCU --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "jps_timed/" -c 400 --csv \
     --log-file gpurun_out/r2b_launches_c4_1gpu.csv python bench.py --quick --steps 2 --warmup 3 > gpurun_out/r2b_ncu_launches.log 2>&1
tail -1 gpurun_out/r2b_ncu_launches.log | cut -c1-200
python tools/make_profiles.py --raw gpurun_out/r2b_raw_c4_1gpu.csv --launches gpurun_out/r2b_launches_c4_1gpu.csv --tag r2 --suffix c4_1gpu \
     --title "C4 on ONE GPU (1e9 uniform particles, PCS, 2048^3): painter kernels from tools/paint_profile.py, launch list from bench.py --quick" > /dev/null
cp profiles/r2_ncu_full_summary_c4_1gpu.md profiles/r2_ncu_dram_traffic_c4_1gpu.json profiles/r2_ncu_launches_c4_1gpu.csv profiles/r2_ncu_launch_shares_c4_1gpu.txt gpurun_out/ 2>/dev/null
# bench lines (the traffic file written above is picked up by the default line)
timeout 900 python bench.py > gpurun_out/r2b_bench_c4_1gpu.json 2> gpurun_out/r2b_bench_c4_1gpu.err; tail -1 gpurun_out/r2b_bench_c4_1gpu.json | cut -c1-400
timeout 600 python bench.py --workload c2 > gpurun_out/r2b_bench_c2_1gpu.json 2>/dev/null; tail -1 gpurun_out/r2b_bench_c2_1gpu.json | cut -c1-300
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2b_bench_reference_arm.json 2>/dev/null; tail -1 gpurun_out/r2b_bench_reference_arm.json | cut -c1-300
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
