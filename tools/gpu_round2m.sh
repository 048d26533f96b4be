#!/bin/bash
mkdir -p gpurun_out
run() { # name
timeout -s KILL 200 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_$1.json 2> gpurun_out/bench_$1.err; python -c "
import json; d=json.load(open('gpurun_out/bench_$1.json')); s=d['stages']; print('$1', round(d['value'],3), round(s['clip_ms'],2), round(s['prefill_ms'],2), round(s['decode_ms'],2), round(d['roofline']['frac'],3), d['gpu_launches'])" || tail -3 gpurun_out/bench_$1.err
}
run base
VCL_TC_OPROJ_SLOTS=12 run o12
VCL_TC_OPROJ_SLOTS=10 run o10
VCL_TC_OPROJ_SLOTS=14 run o14
run base_b
