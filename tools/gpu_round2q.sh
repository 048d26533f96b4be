#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/t_all.log 2>&1; echo "== all gpu tests exit $?"; tail -n 6 gpurun_out/t_all.log | cut -c1-300
timeout -s KILL 200 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_x.json 2> gpurun_out/bench_x.err; python -c "
import json; d=json.load(open('gpurun_out/bench_x.json')); s=d['stages']; print(round(d['value'],3), round(s['clip_ms'],2), round(s['prefill_ms'],2), round(s['decode_ms'],2))"
