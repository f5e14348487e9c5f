#!/bin/bash
# round 2, GPU call 2 (1 GPU): full test suite, K1q v2 A/B, ncu --set full of K1q, bench lines
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r2c2_pytest.log; tail -5 gpurun_out/r2c2_pytest.log
ab() { # kernel chunks env
  env $3 python tools/render_once.py --workload c5 --kernel $1 --chunks $2 --reps 4 2>&1 | tail -3 | python -c "
import sys, json
r=[json.loads(l) for l in sys.stdin]
print('$1 chunks=$2 $3', [round(x['kernel_ms'],3) for x in r], r[-1]['rays_processed'])"
}
{
ab thread 0 X=1
ab pool 0 ZRT_POOL_SLOTS=128
for c in 8 16 32; do ab pool $c ZRT_POOL_SLOTS=64; ab pool $c ZRT_POOL_SLOTS=128; done
} 2>&1 | tee gpurun_out/r2c2_ab.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r2c2_bench_c5_thread.json 2> gpurun_out/r2c2_bench_c5_thread.err; echo "bench thread rc=$?"
python bench.py --steps 5 --warmup 3 --kernel pool --no-cpu --no-configs > gpurun_out/r2c2_bench_c5_pool.json 2> gpurun_out/r2c2_bench_c5_pool.err; echo "bench pool rc=$?"
ZRT_POOL_SLOTS=128 ncu --set full --clock-control none --import-source on -k regex:k_trace_pool -c 1 -f -o gpurun_out/r2c2_prof_pool128 python tools/render_once.py --workload c5 --spp 100 --kernel pool --reps 1 > gpurun_out/r2c2_ncu_pool128.log 2>&1; echo "ncu full rc=$?"
cut -c1-400 gpurun_out/r2c2_bench_c5_thread.json; echo; cut -c1-400 gpurun_out/r2c2_bench_c5_pool.json; tail -3 gpurun_out/r2c2_bench_c5_thread.err
