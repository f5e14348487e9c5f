#!/bin/bash
# round 2, GPU call 17 (1 GPU): K1q two-part items - (slices, parts) grid against the cost model's choice
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/r2c17_pytest.log
ab() { env $5 timeout 200 python tools/render_once.py --workload c5 --kernel pool --chunks 0 --spp $1 --size $2 --reps 3 2>&1 | tail -2 | python -c "
import sys, json
r=[json.loads(l) for l in sys.stdin]
print('c5 spp=$1 size=$2 $5', [round(x['kernel_ms'],3) for x in r], r[-1]['rays_processed'])"; }
{
for spp in 1000 125; do for lk in 32,1 16,1 16,2 8,2 4,2 2,2; do ab $spp 1000 x x ZRT_POOL_LK=$lk; done; ab $spp 1000 x x AUTO=1; done
for spp in 500 250; do for lk in 32,1 16,2 8,2 4,2; do ab $spp 1000 x x ZRT_POOL_LK=$lk; done; ab $spp 1000 x x AUTO=1; done
for size in 500 2000; do for lk in 32,1 8,1 8,2 4,2 2,2; do ab 250 $size x x ZRT_POOL_LK=$lk; done; ab 250 $size x x AUTO=1; done
} 2>&1 | tee gpurun_out/r2c17_ab.log
