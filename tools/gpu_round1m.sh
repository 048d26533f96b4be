#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 240 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "cluster_multicast" -p no:cacheprovider > gpurun_out/t_cl.log 2>&1; echo "== cluster tests exit $?"; tail -n 14 gpurun_out/t_cl.log | cut -c1-500
timeout -s KILL 400 python tools/sweep_gemm.py 2>&1 | tail -12
