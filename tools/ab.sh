#!/bin/bash
# tools/ab.sh NAME... : on the GPU box, run the C2 bench (kernel times) with each tools/variants/NAME.so
cd "$(dirname "$0")/.."
cp jax_powspec_b200/libjps.so /tmp/libjps_orig.so
for v in "$@"; do
  cp tools/variants/$v.so jax_powspec_b200/libjps.so
  echo "== $v"
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --quick-kernels 2>&1 | tail -1 | python -c "
import json,sys
l=json.loads(sys.stdin.read()); print(round(l['ms_per_step'],3), ' '.join(f'{k}={v[\"ms_per_launch\"]:.3f}' for k,v in l['kernels'].items() if k.startswith(('bucket','paint'))))"
  if [ -n "$AB_SLAB" ]; then timeout 200 python tools/slab_rank_paint_profile.py --order 4 2>&1 | tail -1 | cut -c40-300; fi
done
cp /tmp/libjps_orig.so jax_powspec_b200/libjps.so
