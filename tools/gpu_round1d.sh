#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -s -p no:cacheprovider > gpurun_out/t_parity.log 2>&1
echo "== parity exit $?"; grep -E "parity\]|passed|failed|Error|error|assert" gpurun_out/t_parity.log | head -60; tail -n 30 gpurun_out/t_parity.log
