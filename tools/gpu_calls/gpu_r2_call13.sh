#!/bin/bash
set -u
mkdir -p gpurun_out
M=smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct
for o in 0 1; do
  ZRT_ROW_ORDER=$o ncu --metrics $M --clock-control none -k regex:k_trace -c 1 --csv --log-file gpurun_out/r2c13_c2_order$o.csv python tools/render_once.py --workload c2 --reps 1 > /dev/null 2>&1
  grep "k_trace" gpurun_out/r2c13_c2_order$o.csv | awk -F'","' '{print "order '$o'", $(NF-2), $NF}' | tr -d '"'
done
