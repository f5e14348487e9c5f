#!/bin/bash
# round 2, GPU call 2 (1 GPU): full test suite, K1q / K1p A/B against the shipped kernels, bench lines, ncu of K1q
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > gpurun_out/r2c2_pytest.log; tail -8 gpurun_out/r2c2_pytest.log
ab() { # workload kernel chunks env
  env $4 timeout 120 python tools/render_once.py --workload $1 --kernel $2 --chunks $3 --reps 4 2>&1 | tail -3 | python -c "
import sys, json
r=[json.loads(l) for l in sys.stdin]
print('$1 $2 chunks=$3 $4', [round(x['kernel_ms'],3) for x in r], r[-1]['rays_processed'])"
}
{
ab c5 thread 0 X=1
ab c5 pool 0 ZRT_POOL_SLOTS=128
ab c5 pool 0 ZRT_POOL_SPLIT=0
for c in 8 16 32; do ab c5 pool $c ZRT_POOL_SLOTS=64; ab c5 pool $c ZRT_POOL_SLOTS=128; done
for w in c2 c3 c4; do
  ab $w thread 0 X=1; ab $w warp 0 X=1
  for s in 64 96 128; do ab $w pool 0 ZRT_POOL_SLOTS=$s; done
  ab $w pool 0 ZRT_POOL_THRESHOLDS=12,2,4,16; ab $w pool 0 ZRT_POOL_THRESHOLDS=12,2,12,16; ab $w pool 0 ZRT_POOL_THRESHOLDS=16,4,8,24; ab $w pool 0 ZRT_POOL_THRESHOLDS=8,2,8,8
done
} 2>&1 | tee gpurun_out/r2c2_ab.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r2c2_bench_c5_thread.json 2> gpurun_out/r2c2_bench_c5_thread.err; echo "bench thread rc=$?"
python bench.py --steps 5 --warmup 3 --kernel pool --no-cpu --no-configs > gpurun_out/r2c2_bench_c5_pool.json 2> gpurun_out/r2c2_bench_c5_pool.err; echo "bench pool rc=$?"
ZRT_POOL_SLOTS=128 ncu --set full --clock-control none --import-source on -k regex:k_trace_pool -c 1 -f -o gpurun_out/r2c2_prof_pool128 python tools/render_once.py --workload c5 --spp 100 --kernel pool --reps 1 > gpurun_out/r2c2_ncu_pool128.log 2>&1; echo "ncu full rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_trace_bpool -c 1 -f -o gpurun_out/r2c2_prof_bpool_c2 python tools/render_once.py --workload c2 --spp 64 --kernel pool --reps 1 > gpurun_out/r2c2_ncu_bpool_c2.log 2>&1; echo "ncu bpool c2 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_trace_bpool -c 1 -f -o gpurun_out/r2c2_prof_bpool_c4 python tools/render_once.py --workload c4 --spp 16 --kernel pool --reps 1 > gpurun_out/r2c2_ncu_bpool_c4.log 2>&1; echo "ncu bpool c4 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_trace_ws -c 1 -f -o gpurun_out/r2c2_prof_ws_c4 python tools/render_once.py --workload c4 --spp 16 --kernel warp --reps 1 > gpurun_out/r2c2_ncu_ws_c4.log 2>&1; echo "ncu ws c4 rc=$?"
cut -c1-400 gpurun_out/r2c2_bench_c5_thread.json; echo; cut -c1-400 gpurun_out/r2c2_bench_c5_pool.json; tail -3 gpurun_out/r2c2_bench_c5_thread.err
