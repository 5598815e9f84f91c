#!/bin/bash
# Round-2b probe 3 (ONE GPU): bank-class order of a tile's particles in the deposit against arrival order.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
LOG=gpurun_out/r2b_probe3.log
: > $LOG
run1() { echo "== paint n=$2 np=$3 order=$4 [$1]" | tee -a $LOG; env $1 timeout 300 python tools/paint_profile.py --n-mesh $2 --n-part $3 --order $4 2>&1 | tail -1 | tee -a $LOG; }
runr() { echo "== rank$2 [$1]" | tee -a $LOG; env $1 timeout 300 python tools/slab_rank_paint_profile.py --world $2 2>&1 | tail -1 | tee -a $LOG; }
runb() { echo "== bench $2 [$1]" | tee -a $LOG; env $1 timeout 400 python bench.py $2 --quick-kernels --steps 5 2>&1 | tail -1 | python -c "
import json,sys
l=json.loads(sys.stdin.read()); print(round(l['ms_per_step'],3), ' '.join(f'{k}={v[\"ms_per_launch\"]:.3f}' for k,v in l['kernels'].items()))" | tee -a $LOG; }
timeout 900 python -m pytest tests/test_gpu_paint.py tests/test_gpu_slab.py -m gpu -x -q 2>&1 | tail -3 | tee -a $LOG
SAN=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck; do
  timeout 600 $SAN --tool $tool --error-exitcode 99 --print-limit 20 python tools/sanitize_paint.py > gpurun_out/r2b_sanitize_paint_$tool.log 2>&1
  echo "$tool exit code $?" | tee -a gpurun_out/r2b_sanitize_paint_$tool.log
  tail -4 gpurun_out/r2b_sanitize_paint_$tool.log | tee -a $LOG
done
runr "" 8
runr "JPS_TILE_ORDER=arrival" 8
runr "JPS_TILE_THREADS=384" 8
runr "JPS_TILE_THREADS=256" 8
runb "" "--workload c2"
runb "JPS_TILE_ORDER=arrival" "--workload c2"
runb "JPS_TILE_THREADS=512" "--workload c2"
run1 "" 512 1e7 2
run1 "JPS_TILE_ORDER=arrival" 512 1e7 2
run1 "" 2048 1e9 4
run1 "JPS_TILE_ORDER=arrival" 2048 1e9 4
