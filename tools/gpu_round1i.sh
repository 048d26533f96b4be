#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k attention_vit -p no:cacheprovider > gpurun_out/t_attn_tc.log 2>&1; echo "== attn_tc exit $?"; tail -n 30 gpurun_out/t_attn_tc.log | cut -c1-400
