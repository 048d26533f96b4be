#!/bin/bash
mkdir -p gpurun_out
run() { # name
timeout -s KILL 300 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_$1.json 2> gpurun_out/bench_$1.err; python -c "
import json; d=json.load(open('gpurun_out/bench_$1.json')); s=d['stages']; print('$1', round(d['value'],3), round(s['clip_ms'],2), round(s['prefill_ms'],2), round(s['decode_ms'],2), round(d['roofline']['frac'],3), d['gpu_launches'])" || tail -3 gpurun_out/bench_$1.err
}
run tc113
VCL_GEMV_TC_SMEM_KB=150 run tc150
VCL_GEMV_TC_SMEM_KB=190 run tc190
VCL_GEMV_TC_SMEM_KB=226 run tc226
VCL_GEMV_LEGACY=1 run legacy
