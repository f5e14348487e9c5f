#!/bin/bash
# round 2, GPU call 1: new parity tests, K1 vs K1q A/B on c5, quick ncu counters of both
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r2c1_pytest.log; tail -5 gpurun_out/r2c1_pytest.log
ab() { # kernel chunks env
  env $3 python tools/render_once.py --workload c5 --kernel $1 --chunks $2 --reps 4 2>&1 | tail -3 | python -c "
import sys, json
r=[json.loads(l) for l in sys.stdin]
print('$1 chunks=$2 $3', [round(x['kernel_ms'],3) for x in r], r[-1]['rays_processed'])"
}
{
ab thread 0 X=1
for c in 4 8 16 32; do ab pool $c ZRT_POOL_SLOTS=64; ab pool $c ZRT_POOL_SLOTS=128; done
} 2>&1 | tee gpurun_out/r2c1_ab.log
M=smsp__inst_executed.sum,smsp__thread_inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active
for k in thread pool; do
  ncu --metrics $M --clock-control none -k regex:k_trace -c 1 --csv --log-file gpurun_out/r2c1_ncu_$k.csv python tools/render_once.py --workload c5 --spp 100 --kernel $k --reps 1 > gpurun_out/r2c1_ncu_$k.log 2>&1; echo "ncu $k rc=$?"
done
ZRT_POOL_SLOTS=128 ncu --metrics $M --clock-control none -k regex:k_trace -c 1 --csv --log-file gpurun_out/r2c1_ncu_pool128.csv python tools/render_once.py --workload c5 --spp 100 --kernel pool --reps 1 > gpurun_out/r2c1_ncu_pool128.log 2>&1
grep -h "k_trace" gpurun_out/r2c1_ncu_*.csv | cut -d, -f5,13- | head -40
