#!/bin/bash
set -u
mkdir -p gpurun_out
{
for w in c2 c3 c4 c5 c1; do for o in 0 1 2 3; do env ZRT_ROW_ORDER=$o python tools/render_once.py --workload $w --reps 4 2>&1 | tail -3 | python -c "
import sys, json
r=[json.loads(l) for l in sys.stdin]
print('$w row_order=$o', [round(x['kernel_ms'],3) for x in r])"; done; done
for o in 0 1 2 3; do env ZRT_ROW_ORDER=$o python tools/render_once.py --workload c5 --spp 125 --reps 4 2>&1 | tail -3 | python -c "
import sys, json
r=[json.loads(l) for l in sys.stdin]
print('c5@125spp row_order=$o', [round(x['kernel_ms'],3) for x in r])"; done
} 2>&1 | tee gpurun_out/r2c12_ab.log
