#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/t_all.log 2>&1; echo "== all gpu tests exit $?"; tail -n 25 gpurun_out/t_all.log
timeout -s KILL 900 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench2.json 2> gpurun_out/bench2.err; echo "== bench exit $?"; tail -n 5 gpurun_out/bench2.err; python -c "
import json; d=json.load(open('gpurun_out/bench2.json')); print(d['value'], d['e2e']['value'], d['stages'], d['roofline']['frac'], d['gpu_launches'])"
