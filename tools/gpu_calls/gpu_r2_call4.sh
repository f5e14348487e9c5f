#!/bin/bash
# round 2, GPU call 4 (1 GPU): K1q as the default sphere kernel - per-GPU shares of the 8-GPU split, c1, slice counts
set -u
mkdir -p gpurun_out
ab() { # workload kernel chunks spp env
  env $5 timeout 120 python tools/render_once.py --workload $1 --kernel $2 --chunks $3 --spp $4 --reps 4 2>&1 | tail -3 | python -c "
import sys, json
r=[json.loads(l) for l in sys.stdin]
print('$1 $2 chunks=$3 spp=$4 $5', [round(x['kernel_ms'],3) for x in r], r[-1]['rays_processed'])"
}
{
for spp in 125 250 500 1000; do ab c5 thread 0 $spp X=1; ab c5 pool 0 $spp X=1; done
for c in 2 4 8 16; do ab c5 pool $c 125 X=1; done
ab c1 thread 0 100 X=1; ab c1 pool 0 100 X=1; ab c1 pool 8 100 X=1; ab c1 pool 32 100 ZRT_POOL_SLOTS=64
for w in c2 c3 c4; do ab $w pool 0 0 ZRT_POOL_SLOTS=96; ab $w pool 0 0 ZRT_POOL_THRESHOLDS=12,2,8,16,0; ab $w pool 0 0 ZRT_POOL_THRESHOLDS=16,2,12,16,0; done
} 2>&1 | tee gpurun_out/r2c4_ab.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2c4_bench_c5.json 2> gpurun_out/r2c4_bench_c5.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r2c4_bench_c5.json
