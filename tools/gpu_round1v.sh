#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -s -k "336px" -p no:cacheprovider > gpurun_out/t_336.log 2>&1; echo "== 336 test exit $?"; grep -E "parity\]|passed|failed|Error|assert" gpurun_out/t_336.log | head -20 | cut -c1-300
