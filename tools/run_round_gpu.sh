python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c5_v6.json 2> gpurun_out/bench_c5_v6.err; echo "bench rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c5_v6.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_launches_v6.log 2>&1; echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_trace -c 1 -f -o gpurun_out/prof_c5_v6 python tools/render_once.py --workload c5 --reps 1 > gpurun_out/ncu_c5_v6.log 2>&1; echo "ncu full rc=$?"
cut -c1-300 gpurun_out/bench_c5_v6.json
