#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "attention_vit or gemv" -p no:cacheprovider > gpurun_out/t_k.log 2>&1; echo "== tests exit $?"; tail -n 12 gpurun_out/t_k.log | cut -c1-400
timeout -s KILL 300 python tools/microbench.py attn 2>&1 | tail -4
timeout -s KILL 900 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench3.json 2> gpurun_out/bench3.err; echo "== bench exit $?"; tail -n 5 gpurun_out/bench3.err; python -c "
import json; d=json.load(open('gpurun_out/bench3.json')); print(d['value'], d['e2e']['value'], d['stages'], d['roofline']['frac'], d['gpu_launches'])"
