#!/bin/bash
# round 2, GPU call 24 (1 GPU): queue windows larger than one pixel for the BVH kernels and K1
set -u
mkdir -p gpurun_out
{
for w in c2 c3 c4; do for win in 32 64 128 256; do env ZRT_QUEUE_WINDOW_ALL=$win python tools/render_once.py --workload $w --reps 3 2>&1 | tail -2 | python -c "
import sys, json
r=[json.loads(l) for l in sys.stdin]
print('$w window=$win', [round(x['kernel_ms'],3) for x in r])"; done; done
for win in 32 128; do env ZRT_QUEUE_WINDOW_ALL=$win python tools/render_once.py --workload c5 --kernel thread --reps 3 2>&1 | tail -2 | python -c "
import sys, json
r=[json.loads(l) for l in sys.stdin]
print('c5 thread window=$win', [round(x['kernel_ms'],3) for x in r])"; done
} 2>&1 | tee gpurun_out/r2c24_ab.log
