#!/bin/bash
set -u
mkdir -p gpurun_out
ab() { env $5 timeout 120 python tools/render_once.py --workload $1 --kernel $2 --chunks $3 --spp $4 --reps 4 2>&1 | tail -3 | python -c "
import sys, json
r=[json.loads(l) for l in sys.stdin]
print('$1 $2 chunks=$3 spp=$4 $5', [round(x['kernel_ms'],3) for x in r], r[-1]['rays_processed'])"; }
{
for c in 32 8; do for spp in 1000 125; do ab c5 pool $c $spp ZRT_ROWS_TOP_DOWN=0; ab c5 pool $c $spp ZRT_ROWS_TOP_DOWN=1; done; done
ab c5 thread 8 1000 ZRT_ROWS_TOP_DOWN=0; ab c5 thread 8 1000 ZRT_ROWS_TOP_DOWN=1
} 2>&1 | tee gpurun_out/r2c10_ab.log
