#!/bin/bash
# first GPU pass: per-kernel parity (one process per kernel family so a hang cannot mask the rest)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for k in st_pool layernorm gemv attention gemm; do
  timeout -s KILL 420 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k $k -p no:cacheprovider > gpurun_out/t_$k.log 2>&1
  echo "== $k exit $?"; tail -n 25 gpurun_out/t_$k.log
done
timeout -s KILL 420 python tools/microbench.py > gpurun_out/micro.log 2>&1
echo "== micro exit $?"; cat gpurun_out/micro.log
