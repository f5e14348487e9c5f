#!/bin/bash
# bash tools/ab_kernels.sh "<variants>" "<workloads>" "<kernels: auto warp ...>"
for v in $1; do
  cp variants/libzrt_$v.so zraytrace_b200/libzrt.so
  for w in $2; do for k in $3; do
    python tools/render_once.py --workload $w --kernel $k --reps 5 2>&1 | tail -4 | python -c "
import sys, json
print('$v $w $k', [round(json.loads(l)['kernel_ms'], 3) for l in sys.stdin])"
  done; done
done
