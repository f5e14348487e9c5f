#!/bin/bash
# One-box measurement pass of the round: GPU tests, default bench line, ncu launch list + full capture of the headline
# kernel (each only after its plain command exited 0), the reference arm, the other workloads.
#   gpurun --timeout 1200 -- 'bash tools/run_round_gpu.sh v10'
set -u
V=${1:-v10}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c5_$V.json 2> gpurun_out/bench_c5_$V.err; echo "bench c5 rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_c5_$V.json 2> gpurun_out/bench_ref_c5_$V.err; echo "bench reference rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c5_$V.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_launches_$V.log 2>&1; echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_trace -c 1 -f -o gpurun_out/prof_c5_$V python tools/render_once.py --workload c5 --reps 1 > gpurun_out/ncu_c5_$V.log 2>&1; echo "ncu full rc=$?"
for w in c1 c2 c3 c4; do python bench.py --workload $w --steps 3 --warmup 3 > gpurun_out/bench_${w}_$V.json 2> gpurun_out/bench_${w}_$V.err; echo "bench $w rc=$?"; done
cut -c1-200 gpurun_out/bench_c5_$V.json; cut -c1-200 gpurun_out/bench_ref_c5_$V.json
