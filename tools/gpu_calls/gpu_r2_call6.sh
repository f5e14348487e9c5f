#!/bin/bash
# round 2, GPU call 6 (1 GPU): K1q with the hand-over ring against the previous build (tools/ab/libzrt_prev.so), tests, bench, ncu
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/r2c6_pytest.log
ab() { # workload kernel chunks spp env
  env $5 timeout 120 python tools/render_once.py --workload $1 --kernel $2 --chunks $3 --spp $4 --reps 4 2>&1 | tail -3 | python -c "
import sys, json
r=[json.loads(l) for l in sys.stdin]
print('$1 $2 chunks=$3 spp=$4 $5', [round(x['kernel_ms'],3) for x in r], r[-1]['rays_processed'])"
}
{
for spp in 1000 125; do ab c5 pool 0 $spp ZRT_LIB_PATH=$PWD/tools/ab/libzrt_prev.so; ab c5 pool 0 $spp X=1; done
ab c5 pool 16 1000 X=1; ab c5 pool 8 1000 X=1; ab c5 pool 16 125 X=1
ab c1 pool 0 100 X=1; ab c1 thread 0 100 X=1
} 2>&1 | tee gpurun_out/r2c6_ab.log
ncu --set full --clock-control none --import-source on -k regex:k_trace_pool3 -c 1 -f -o gpurun_out/r2c6_prof_pool3 python tools/render_once.py --workload c5 --spp 1000 --kernel pool --reps 1 > gpurun_out/r2c6_ncu_pool3.log 2>&1; echo "ncu full rc=$?"
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2c6_bench_c5.json 2> gpurun_out/r2c6_bench_c5.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r2c6_bench_c5.json
