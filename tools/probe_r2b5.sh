#!/bin/bash
# Round-2b probe 5 (ONE GPU): PCS deposit = bank-class order + fixed point at 2^-28 + carry count by add-with-carry,
# against arrival order / 2^-31; TSC and CIC must be back to their arrival-order times.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
LOG=gpurun_out/r2b_probe5.log
: > $LOG
run1() { echo "== paint n=$2 np=$3 order=$4 [$1]" | tee -a $LOG; env $1 timeout 300 python tools/paint_profile.py --n-mesh $2 --n-part $3 --order $4 2>&1 | tail -1 | tee -a $LOG; }
runr() { echo "== rank$2 [$1]" | tee -a $LOG; env $1 timeout 300 python tools/slab_rank_paint_profile.py --world $2 2>&1 | tail -1 | tee -a $LOG; }
runb() { echo "== bench $2 [$1]" | tee -a $LOG; env $1 timeout 400 python bench.py $2 --quick-kernels --steps 5 2>&1 | tail -1 | python -c "
import json,sys
l=json.loads(sys.stdin.read()); print(round(l['ms_per_step'],3), ' '.join(f'{k}={v[\"ms_per_launch\"]:.3f}' for k,v in l['kernels'].items()))" | tee -a $LOG; }
timeout 900 python -m pytest tests/test_gpu_paint.py -m gpu -x -q 2>&1 | tail -3 | tee -a $LOG
SAN=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck; do
  timeout 600 $SAN --tool $tool --error-exitcode 99 --print-limit 20 python tools/sanitize_paint.py > gpurun_out/r2b_sanitize_paint_$tool.log 2>&1
  echo "$tool exit code $?" | tee -a gpurun_out/r2b_sanitize_paint_$tool.log
  tail -3 gpurun_out/r2b_sanitize_paint_$tool.log | tee -a $LOG
done
runr "" 8
runr "JPS_FX_BITS=31" 8
runr "JPS_FX_BITS=29" 8
runr "JPS_TILE_ORDER=arrival" 8
runr "JPS_TILE_ORDER=arrival JPS_FX_BITS=28" 8
runr "JPS_TILE_THREADS=384" 8
runb "" "--workload c2"
runb "JPS_TILE_ORDER=bank JPS_FX_BITS=28" "--workload c2"
run1 "" 512 1e7 2
run1 "" 2048 1e9 4
run1 "JPS_TILE_ORDER=arrival" 2048 1e9 4
NCU=/usr/local/cuda/bin/ncu
M=l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__inst_executed.sum,gpu__time_duration.sum,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,sm__warps_active.avg.pct_of_peak_sustained_active
echo "== rank8 deposit ncu (default)" | tee -a $LOG
timeout 300 $NCU --metrics $M --clock-control none -k regex:paint_tile_fx --launch-skip 3 --launch-count 1 python tools/slab_rank_paint_profile.py 2>&1 | grep -E "l1tex|smsp|sm__|gpu__time" | tee -a $LOG
