#!/bin/bash
# round 2, GPU call 21 (1 GPU): K1q queue window, finer grid with repeats
set -u
mkdir -p gpurun_out
ab() { env $4 timeout 200 python tools/render_once.py --workload c5 --kernel pool --chunks $1 --spp $2 --size $3 --reps 3 2>&1 | tail -2 | python -c "
import sys, json
r=[json.loads(l) for l in sys.stdin]
print('c5 chunks=$1 spp=$2 size=$3 $4', [round(x['kernel_ms'],3) for x in r])"; }
{
for rep in 1 2; do for w in 32 64 96 128 160 192 224 256 320 384; do ab 32 1000 1000 ZRT_QUEUE_WINDOW=$w; done; done
for w in 32 64 128 192; do ab 16 125 1000 ZRT_QUEUE_WINDOW=$w; ab 32 500 1000 ZRT_QUEUE_WINDOW=$w; done
} 2>&1 | tee gpurun_out/r2c21_ab.log
