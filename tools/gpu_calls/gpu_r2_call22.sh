#!/bin/bash
# round 2, GPU call 22 (1 GPU): K1q slices x queue window with the window rule in place; parity suite on the shipped default
set -u
mkdir -p gpurun_out
ab() { env $4 timeout 200 python tools/render_once.py --workload c5 --kernel pool --chunks $1 --spp $2 --size $3 --reps 3 2>&1 | tail -2 | python -c "
import sys, json
r=[json.loads(l) for l in sys.stdin]
print('c5 chunks=$1 spp=$2 size=$3 $4', [round(x['kernel_ms'],3) for x in r])"; }
{
for spp in 1000 125; do ab 0 $spp 1000 AUTO=1; ab 32 $spp 1000 ZRT_QUEUE_WINDOW=128; ab 16 $spp 1000 ZRT_QUEUE_WINDOW=64; ab 16 $spp 1000 ZRT_QUEUE_WINDOW=128; ab 8 $spp 1000 ZRT_QUEUE_WINDOW=32; ab 8 $spp 1000 ZRT_QUEUE_WINDOW=64; done
ab 0 250 2000 AUTO=1; ab 8 250 2000 ZRT_QUEUE_WINDOW=32; ab 8 250 2000 ZRT_QUEUE_WINDOW=64; ab 0 250 500 AUTO=1
} 2>&1 | tee gpurun_out/r2c22_ab.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -2
