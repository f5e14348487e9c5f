#!/bin/bash
# round 2, GPU call 20 (1 GPU): ncu --set full of the FINAL build's dominant kernels (c5 K1q at 1000 spp, c2/c3/c4 K1w at their full sample counts where affordable)
set -u
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_trace_pool3 -c 1 -f -o gpurun_out/r2w_prof_c5_pool3 python tools/render_once.py --workload c5 --reps 1 > gpurun_out/r2w_ncu_c5.log 2>&1; echo "ncu c5 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_trace_ws -c 1 -f -o gpurun_out/r2w_prof_c2_ws python tools/render_once.py --workload c2 --reps 1 > gpurun_out/r2w_ncu_c2.log 2>&1; echo "ncu c2 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_trace_ws -c 1 -f -o gpurun_out/r2w_prof_c3_ws python tools/render_once.py --workload c3 --spp 128 --reps 1 > gpurun_out/r2w_ncu_c3.log 2>&1; echo "ncu c3 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_trace_ws -c 1 -f -o gpurun_out/r2w_prof_c4_ws python tools/render_once.py --workload c4 --spp 128 --reps 1 > gpurun_out/r2w_ncu_c4.log 2>&1; echo "ncu c4 rc=$?"
ls -la gpurun_out/r2w_*.ncu-rep
