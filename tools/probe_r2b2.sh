#!/bin/bash
# Round-2b probe 2 (ONE GPU): auto rule of the staged fine pass, tile histogram fused into the coarse pass, 8192-record
# coarse chunks; on the painter alone (full 2048^3 mesh, one rank of the 8-GPU and of the 2-GPU decomposition), C2, and
# the default bench's per-kernel pass.
# Variant library (not kept in the tree; build it first):
#   bash tools/build_variant.sh coarse8k -DJPS_COARSE_CHUNK=8192 -DJPS_COARSE_THREADS=1024 -DJPS_COARSE_MINB=1
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
LOG=gpurun_out/r2b_probe2.log
: > $LOG
cp jax_powspec_b200/libjps.so /tmp/libjps_base.so
use() { cp "$1" jax_powspec_b200/libjps.so; }
run1() { echo "== 1gpu [$1] lib=$2" | tee -a $LOG; env $1 timeout 300 python tools/paint_profile.py --tag "$2 $1" 2>&1 | tail -1 | tee -a $LOG; }
runr() { echo "== rank$3 [$1] lib=$2" | tee -a $LOG; env $1 timeout 300 python tools/slab_rank_paint_profile.py --world $3 2>&1 | tail -1 | tee -a $LOG; }
runb() { echo "== bench $3 [$1] lib=$2" | tee -a $LOG; env $1 timeout 400 python bench.py $3 --quick-kernels --steps 3 2>&1 | tail -1 | python -c "
import json,sys
l=json.loads(sys.stdin.read()); print(round(l['ms_per_step'],3), ' '.join(f'{k}={v[\"ms_per_launch\"]:.3f}' for k,v in l['kernels'].items()), 'check', l.get('check',{}).get('ok'))" | tee -a $LOG; }

timeout 900 python -m pytest tests/test_gpu_paint.py tests/test_gpu_slab.py -m gpu -x -q 2>&1 | tail -3 | tee -a $LOG

run1 "" base
run1 "JPS_COUNT=fused" base
runr "" base 8
runr "JPS_FINE_CHUNK=big" base 8
runr "JPS_COUNT=fused" base 8
runr "" base 2
runr "JPS_COUNT=fused" base 2
runb "" base "--workload c2"
runb "" base ""
runb "JPS_COUNT=fused" base ""
use tools/variants/coarse8k.so
run1 "" coarse8k
run1 "JPS_COUNT=fused" coarse8k
runr "" coarse8k 8
runb "" coarse8k "--workload c2"
use /tmp/libjps_base.so
