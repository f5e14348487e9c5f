#!/bin/bash
# round 2, measurement pass (1 GPU): what the driver runs at round end, plus the launch list and the other workloads
set -u
V=${1:-r2_g}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/${V}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/${V}_smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${V}_bench_c5.json 2> gpurun_out/${V}_bench_c5.err; echo "bench c5 rc=$?"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${V}_bench_ref_c5.json 2> gpurun_out/${V}_bench_ref_c5.err; echo "bench reference rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${V}_launches_c5.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-configs > gpurun_out/${V}_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
for w in c1 c2 c3 c4; do timeout 600 python bench.py --workload $w --steps 3 --warmup 3 > gpurun_out/${V}_bench_$w.json 2> gpurun_out/${V}_bench_$w.err; echo "bench $w rc=$?"; done
cut -c1-250 gpurun_out/${V}_bench_c5.json; cut -c1-250 gpurun_out/${V}_bench_ref_c5.json
