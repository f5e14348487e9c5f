#!/bin/bash
set -u
mkdir -p gpurun_out
{
for w in c4 c2 c3 c5; do for t in 0 3 4 5 6; do env ZRT_TILE_LOG2=$t python tools/render_once.py --workload $w --reps 3 2>&1 | tail -2 | python -c "
import sys, json
r=[json.loads(l) for l in sys.stdin]
print('$w tile_log2=$t', [round(x['kernel_ms'],3) for x in r])"; done; done
} 2>&1 | tee gpurun_out/r2c16_ab.log
