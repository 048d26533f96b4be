#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=8 > gpurun_out/t_all.log 2>&1; echo "== all gpu tests exit $?"; tail -n 16 gpurun_out/t_all.log
for cl in 0 1 0 1; do
VCL_GEMM_CLUSTER=$cl timeout -s KILL 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_cl$cl.json 2> gpurun_out/bench_cl$cl.err; python -c "
import json; d=json.load(open('gpurun_out/bench_cl$cl.json')); s=d['stages']; print('cl=$cl', round(d['value'],3), round(s['clip_ms'],2), round(s['prefill_ms'],2), round(s['decode_ms'],2), d['clocks'])"
done
