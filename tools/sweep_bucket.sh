#!/bin/bash
# parameter sweep of the bucketing passes (run on the GPU box)
for rep in 1 2 4 8; do for unr in 1 2 4; do for bp in 8 16 32; do
  JPS_BUCKET_REP=$rep JPS_SCATTER_UNROLL=$unr JPS_SCATTER_BLOCKS_PER_SM=$bp timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu --quick-kernels 2>&1 | tail -1 | python -c "
import json,sys
l=json.loads(sys.stdin.read()); k=l['kernels']; print('rep=$rep unr=$unr bpsm=$bp', 'step', round(l['ms_per_step'],3), 'count', round(k['bucket_count']['ms_per_launch'],3), 'scatter', round(k['bucket_scatter']['ms_per_launch'],3), 'tile', round(k['paint_tile']['ms_per_launch'],3))"
done; done; done
