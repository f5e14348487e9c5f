#!/bin/bash
# round 2, GPU call 7 (1 GPU): K1q hand-over policy (inline above a threshold, parked below) over the per-GPU shares
set -u
mkdir -p gpurun_out
ab() { # workload kernel chunks spp env
  env $5 timeout 120 python tools/render_once.py --workload $1 --kernel $2 --chunks $3 --spp $4 --reps 4 2>&1 | tail -3 | python -c "
import sys, json
r=[json.loads(l) for l in sys.stdin]
print('$1 $2 chunks=$3 spp=$4 $5', [round(x['kernel_ms'],3) for x in r], r[-1]['rays_processed'])"
}
{
for spp in 1000 500 250 125; do for hm in 1 4 6 10 33; do ab c5 pool 0 $spp ZRT_POOL_HAND_MIN=$hm; done; done
ab c5 pool 16 125 ZRT_POOL_HAND_MIN=6; ab c5 pool 16 250 ZRT_POOL_HAND_MIN=6
ab c1 pool 0 100 X=1; ab c1 thread 0 100 X=1
} 2>&1 | tee gpurun_out/r2c7_ab.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
