#!/bin/bash
set -u
mkdir -p gpurun_out
ab() { env $6 timeout 200 python tools/render_once.py --workload $1 --kernel $2 --chunks $3 --spp $4 --size $5 --reps 3 2>&1 | tail -2 | python -c "
import sys, json
r=[json.loads(l) for l in sys.stdin]
print('$1 $2 chunks=$3 spp=$4 size=$5 $6', [round(x['kernel_ms'],3) for x in r], r[-1]['rays_processed'])"; }
{
for size in 500 1000 2000; do for c in 32 8; do ab c5 pool $c 250 $size X=1; done; done
for w in c2 c3 c4; do for o in 0 1; do env ZRT_ROWS_TOP_DOWN=$o python tools/render_once.py --workload $w --reps 4 2>&1 | tail -3 | python -c "
import sys, json
r=[json.loads(l) for l in sys.stdin]
print('$w rows_top_down=$o', [round(x['kernel_ms'],3) for x in r])"; done; done
} 2>&1 | tee gpurun_out/r2c11_ab.log
