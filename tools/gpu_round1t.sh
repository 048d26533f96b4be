#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python bench.py --clips 16 --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_b16.json 2> gpurun_out/bench_b16.err; tail -n 3 gpurun_out/bench_b16.err; python -c "
import json; d=json.load(open('gpurun_out/bench_b16.json')); s=d['stages']; print('B=16', round(d['value'],3), round(s['clip_ms'],2), round(s['prefill_ms'],2), round(s['decode_ms'],2), round(s['clip_frac'],3), round(s['prefill_frac'],3), d['gpu_launches'])"
timeout -s KILL 900 python bench.py --model 13b --clips 4 --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_13b.json 2> gpurun_out/bench_13b.err; tail -n 3 gpurun_out/bench_13b.err; python -c "
import json; d=json.load(open('gpurun_out/bench_13b.json')); s=d['stages']; print('13B B=4', round(d['value'],3), round(s['clip_ms'],2), round(s['prefill_ms'],2), round(s['decode_ms'],2), round(d['roofline']['frac'],3), d['gpu_launches'])"
