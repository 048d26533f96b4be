#!/bin/bash
mkdir -p gpurun_out
for k in gemm st_pool; do
  timeout -s KILL 420 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k $k -p no:cacheprovider > gpurun_out/t_$k.log 2>&1
  echo "== $k exit $?"; tail -n 12 gpurun_out/t_$k.log
done
timeout -s KILL 420 python tools/microbench.py gemm > gpurun_out/micro.log 2>&1
echo "== micro exit $?"; cat gpurun_out/micro.log
