#!/bin/bash
# round 2, GPU call 3 (1 GPU): ncu --set full with source of the SHIPPED BVH kernels on c2/c3/c4 (reduced spp, same plane)
set -u
mkdir -p gpurun_out
for cfg in "c2 64 warp" "c3 32 warp" "c3 32 thread" "c4 16 warp"; do
  set -- $cfg
  python tools/render_once.py --workload $1 --spp $2 --kernel $3 --reps 2 2>&1 | tail -1 | cut -c1-200
  ncu --set full --clock-control none --import-source on -k regex:k_trace -c 1 -f -o gpurun_out/r2c3_prof_$1_$3 python tools/render_once.py --workload $1 --spp $2 --kernel $3 --reps 1 > gpurun_out/r2c3_ncu_$1_$3.log 2>&1; echo "ncu $1 $3 rc=$?"
done
ls -la gpurun_out/*.ncu-rep
