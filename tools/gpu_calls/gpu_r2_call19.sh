#!/bin/bash
set -u
mkdir -p gpurun_out
{
for w in c2 c3 c4; do for k in thread warp; do python tools/render_once.py --workload $w --kernel $k --reps 3 2>&1 | tail -2 | python -c "
import sys, json
r=[json.loads(l) for l in sys.stdin]
print('$w $k', [round(x['kernel_ms'],3) for x in r])"; done; done
for th in 12,2,20 12,2,16 16,2,20 8,2,20 12,4,20 12,2,24 12,1,20; do for w in c2 c3 c4; do ZRT_WS_THRESHOLDS=$th python tools/render_once.py --workload $w --kernel warp --reps 3 2>&1 | tail -2 | python -c "
import sys, json
r=[json.loads(l) for l in sys.stdin]
print('$w warp th=$th', [round(x['kernel_ms'],3) for x in r])"; done; done
} 2>&1 | tee gpurun_out/r2c19_ab.log
