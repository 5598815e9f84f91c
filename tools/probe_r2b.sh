#!/bin/bash
# Round-2b probe (ONE GPU): bucketing passes of the painter at C4's full size on one GPU -- group count, staged fine
# pass, streaming loads, fixed-point position of the deposit -- plus a limited-section ncu capture of the three passes.
# Variant libraries (not kept in the tree; build them first):
#   bash tools/build_variant.sh fx28 -DJPS_FX_BITS=28      (before fx_bits became a kernel argument: now JPS_FX_BITS=28 in the environment)
#   bash tools/build_variant.sh ldcs -DJPS_FINE_LDCS
#   bash tools/build_variant.sh fs8k -DJPS_FS_CHUNK=8192 -DJPS_FS_THREADS=1024 -DJPS_FS_MINB=1   (now JPS_FINE_CHUNK=big)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
LOG=gpurun_out/r2b_probe.log
: > $LOG
cp jax_powspec_b200/libjps.so /tmp/libjps_base.so
use() { cp "$1" jax_powspec_b200/libjps.so; }
run1() { echo "== 1gpu [$1] lib=$2" | tee -a $LOG; env $1 timeout 300 python tools/paint_profile.py --tag "$2 $1" 2>&1 | tail -1 | tee -a $LOG; }
runr() { echo "== rank8 [$1] lib=$2" | tee -a $LOG; env $1 timeout 300 python tools/slab_rank_paint_profile.py 2>&1 | tail -1 | tee -a $LOG; }
runc2() { echo "== C2 [$1] lib=$2" | tee -a $LOG; env $1 timeout 300 python bench.py --workload c2 --quick-kernels --steps 5 2>&1 | tail -1 | python -c "
import json,sys
l=json.loads(sys.stdin.read()); print(round(l['ms_per_step'],3), ' '.join(f'{k}={v[\"ms_per_launch\"]:.3f}' for k,v in l['kernels'].items()))" | tee -a $LOG; }

# correctness first: the painter tests (incl. both fine-pass forms in subprocesses)
timeout 600 python -m pytest tests/test_gpu_paint.py -m gpu -x -q 2>&1 | tail -3 | tee -a $LOG

run1 "" base
run1 "JPS_FINE=staged" base
run1 "JPS_MAX_GROUPS=4096" base
run1 "JPS_MAX_GROUPS=4096 JPS_FINE=staged" base
run1 "JPS_MAX_GROUPS=1024 JPS_FINE=staged" base
use tools/variants/ldcs.so; run1 "" ldcs; run1 "JPS_FINE=staged" ldcs
use tools/variants/fs8k.so; run1 "JPS_FINE=staged" fs8k; run1 "JPS_FINE=staged JPS_MAX_GROUPS=4096" fs8k
use tools/variants/fx28.so; run1 "" fx28; runr "" fx28; runc2 "" fx28
use /tmp/libjps_base.so
runr "" base; runr "JPS_FINE=staged" base
runc2 "" base; runc2 "JPS_FINE=staged" base

# ncu: the three bucketing passes at full size, direct and staged fine pass (2 warm-up paints x 3 matching kernels skipped)
NCU=/usr/local/cuda/bin/ncu
SECS="--section SpeedOfLight --section MemoryWorkloadAnalysis --section WarpStateStats --section Occupancy --section SchedulerStats --section LaunchStats"
timeout 600 $NCU $SECS --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none \
   -k regex:'fine_scatter|fine_staged|coarse_scatter|bucket_count' --launch-skip 6 --launch-count 3 -f -o gpurun_out/r2b_prof_c4_1gpu_bucket \
   python tools/paint_profile.py --warmup 2 --reps 1 > gpurun_out/r2b_ncu_direct.log 2>&1
tail -2 gpurun_out/r2b_ncu_direct.log
JPS_FINE=staged timeout 600 $NCU $SECS --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none \
   -k regex:'fine_staged' --launch-skip 2 --launch-count 1 -f -o gpurun_out/r2b_prof_c4_1gpu_fine_staged \
   python tools/paint_profile.py --warmup 2 --reps 1 > gpurun_out/r2b_ncu_staged.log 2>&1
tail -2 gpurun_out/r2b_ncu_staged.log
ls -la gpurun_out/*.ncu-rep
