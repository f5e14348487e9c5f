#!/bin/bash
set -u
mkdir -p gpurun_out
M=smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_elapsed,gpu__time_duration.sum,sm__cycles_active.avg,sm__cycles_elapsed.avg,sm__cycles_active.min,sm__cycles_active.max
for cfg in "c2 256" "c3 128" "c4 128" "c5 250"; do set -- $cfg
  for o in 0 1; do
  ZRT_ROW_ORDER=$o ncu --metrics $M --clock-control none -k regex:k_trace -c 1 --csv --log-file gpurun_out/r2c15_$1_order$o.csv python tools/render_once.py --workload $1 --spp $2 --reps 1 > /dev/null 2>&1
  grep "k_trace" gpurun_out/r2c15_$1_order$o.csv | awk -F'","' '{printf "%s order '$o' %s %s\n", "'$1'", $(NF-2), $NF}' | tr -d '"'
  done
done
