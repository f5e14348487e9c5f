#!/bin/bash
# round 2, multi-GPU call (gpurun --gpus N): the C-ABI group on hardware - tests, single-process and torchrun bench lines
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q -x -s 2>&1 | tail -12 | tee gpurun_out/r2_multi_pytest_n$N.log
for n in 1 2 4 8; do
  [ $n -le $N ] || continue
  timeout 300 python bench.py --gpus $n --steps 5 --warmup 3 --no-cpu --no-configs > gpurun_out/r2_scale_single_n$n.json 2> gpurun_out/r2_scale_single_n$n.err; echo "single-process n=$n rc=$?"; cut -c1-300 gpurun_out/r2_scale_single_n$n.json
  if [ $n -gt 1 ]; then
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 5 --warmup 3 --no-cpu --no-configs > gpurun_out/r2_scale_torchrun_n$n.json 2> gpurun_out/r2_scale_torchrun_n$n.err; echo "torchrun n=$n rc=$?"; cut -c1-300 gpurun_out/r2_scale_torchrun_n$n.json
  fi
done
