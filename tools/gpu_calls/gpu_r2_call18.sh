#!/bin/bash
# round 2, GPU call 18 (1 GPU): K1q queue window (items per atomic) x slices per pixel
set -u
mkdir -p gpurun_out
ab() { env $4 timeout 200 python tools/render_once.py --workload c5 --kernel pool --chunks $1 --spp $2 --size $3 --reps 3 2>&1 | tail -2 | python -c "
import sys, json
r=[json.loads(l) for l in sys.stdin]
print('c5 chunks=$1 spp=$2 size=$3 $4', [round(x['kernel_ms'],3) for x in r], r[-1]['rays_processed'])"; }
{
for w in 32 128 256 512 1024; do ab 32 1000 1000 ZRT_QUEUE_WINDOW=$w; done
for w in 32 256 1024; do ab 16 1000 1000 ZRT_QUEUE_WINDOW=$w; ab 8 1000 1000 ZRT_QUEUE_WINDOW=$w; done
for w in 32 256 1024; do ab 16 125 1000 ZRT_QUEUE_WINDOW=$w; ab 8 125 1000 ZRT_QUEUE_WINDOW=$w; ab 4 125 1000 ZRT_QUEUE_WINDOW=$w; done
for w in 32 256; do ab 32 250 500 ZRT_QUEUE_WINDOW=$w; ab 8 250 2000 ZRT_QUEUE_WINDOW=$w; done
} 2>&1 | tee gpurun_out/r2c18_ab.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -2
