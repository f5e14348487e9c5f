#!/bin/bash
set -u
mkdir -p gpurun_out
{
for spp in 64 256 1024; do for o in 0 1; do env ZRT_ROW_ORDER=$o python tools/render_once.py --workload c2 --spp $spp --reps 3 2>&1 | tail -2 | python -c "
import sys, json
r=[json.loads(l) for l in sys.stdin]
print('c2 spp=$spp row_order=$o', [round(x['kernel_ms'],3) for x in r])"; done; done
for k in thread warp; do for o in 0 1; do env ZRT_ROW_ORDER=$o python tools/render_once.py --workload c2 --kernel $k --reps 3 2>&1 | tail -2 | python -c "
import sys, json
r=[json.loads(l) for l in sys.stdin]
print('c2 kernel=$k row_order=$o', [round(x['kernel_ms'],3) for x in r])"; done; done
for c in 32 16 8; do env ZRT_ROW_ORDER=1 python tools/render_once.py --workload c2 --chunks $c --reps 3 2>&1 | tail -2 | python -c "
import sys, json
r=[json.loads(l) for l in sys.stdin]
print('c2 chunks=$c row_order=1', [round(x['kernel_ms'],3) for x in r])"; done
} 2>&1 | tee gpurun_out/r2c14_ab.log
