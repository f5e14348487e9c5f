#!/bin/bash
# round 2, GPU call 8 (1 GPU): K1q with ended items routed straight to the hand-over ring, against the previous build
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/r2c8_pytest.log
ab() { # workload kernel chunks spp env
  env $5 timeout 120 python tools/render_once.py --workload $1 --kernel $2 --chunks $3 --spp $4 --reps 4 2>&1 | tail -3 | python -c "
import sys, json
r=[json.loads(l) for l in sys.stdin]
print('$1 $2 chunks=$3 spp=$4 $5', [round(x['kernel_ms'],3) for x in r], r[-1]['rays_processed'])"
}
{
for spp in 1000 500 250 125; do for c in 32 16 8; do ab c5 pool $c $spp ZRT_LIB_PATH=$PWD/tools/ab/libzrt_prev.so; ab c5 pool $c $spp X=1; done; done
ab c5 pool 4 125 X=1; ab c5 pool 0 125 X=1; ab c5 pool 0 1000 X=1
ab c1 pool 0 100 X=1; ab c1 pool 8 100 X=1; ab c1 thread 0 100 X=1
} 2>&1 | tee gpurun_out/r2c8_ab.log
