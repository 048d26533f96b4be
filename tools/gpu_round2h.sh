#!/bin/bash
mkdir -p gpurun_out
run() { # name
timeout -s KILL 300 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_$1.json 2> gpurun_out/bench_$1.err; python -c "
import json; d=json.load(open('gpurun_out/bench_$1.json')); s=d['stages']; print('$1', round(d['value'],3), round(s['clip_ms'],2), round(s['prefill_ms'],2), round(s['decode_ms'],2), round(d['roofline']['frac'],3), d['gpu_launches'])" || tail -3 gpurun_out/bench_$1.err
}
for kb in 150 130 140 165 120 150; do VCL_GEMV_TC_SMEM_KB=$kb run tc$kb; done
VCL_GEMV_LEGACY=1 run legacy
VCL_GEMV_TC_SMEM_KB=150 timeout -s KILL 600 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/t_all.log 2>&1; echo "== all gpu tests exit $?"; tail -n 3 gpurun_out/t_all.log | cut -c1-300
