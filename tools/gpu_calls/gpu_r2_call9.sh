#!/bin/bash
# round 2, GPU call 9 (1 GPU): where the per-launch constant of K1q comes from (time against samples per pixel at fixed slices)
set -u
mkdir -p gpurun_out
ab() { # workload kernel chunks spp env
  env $5 timeout 120 python tools/render_once.py --workload $1 --kernel $2 --chunks $3 --spp $4 --reps 4 2>&1 | tail -3 | python -c "
import sys, json
r=[json.loads(l) for l in sys.stdin]
print('$1 $2 chunks=$3 spp=$4 $5', [round(x['kernel_ms'],3) for x in r], r[-1]['rays_processed'])"
}
{
for c in 32 16 8 4; do for spp in 32 64 128 256; do ab c5 pool $c $spp X=1; done; done
for spp in 32 64 128 256; do ab c5 thread 8 $spp X=1; done
} 2>&1 | tee gpurun_out/r2c9_ab.log
