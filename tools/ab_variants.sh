#!/bin/bash
# A/B of prebuilt library variants on one box: bash tools/ab_variants.sh "base s9" "c5 c1" [spp]
# (variants/libzrt_<name>.so are built here and travel with the snapshot; variants/ is git-ignored)
for v in $1; do
  cp variants/libzrt_$v.so zraytrace_b200/libzrt.so
  for w in $2; do
    python tools/render_once.py --workload $w ${3:+--spp $3} --reps 5 2>&1 | tail -4 | python -c "
import sys, json
print('$v $w ${3:-}', [round(json.loads(l)['kernel_ms'], 3) for l in sys.stdin])"
  done
done
