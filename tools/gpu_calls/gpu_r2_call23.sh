#!/bin/bash
# round 2, GPU call 23 (1 GPU): image rings of their own, re-checked on the final kernel
set -u
mkdir -p gpurun_out
ab() { env $3 timeout 200 python tools/render_once.py --workload c5 --kernel pool --chunks 0 --spp $1 --size $2 --reps 3 2>&1 | tail -2 | python -c "
import sys, json
r=[json.loads(l) for l in sys.stdin]
print('c5 spp=$1 size=$2 $3', [round(x['kernel_ms'],3) for x in r])"; }
{
for rep in 1 2; do ab 1000 1000 ZRT_POOL_SPLIT=0; ab 1000 1000 ZRT_POOL_SPLIT=1; done
ab 125 1000 ZRT_POOL_SPLIT=0; ab 125 1000 ZRT_POOL_SPLIT=1
} 2>&1 | tee gpurun_out/r2c23_ab.log
