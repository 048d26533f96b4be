#!/bin/bash
mkdir -p gpurun_out
echo skip-tests
date
SECONDS=0; timeout -s KILL 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench wall ${SECONDS}s"; tail -1 gpurun_out/bench_default.err; python -c "
import json; d=json.load(open('gpurun_out/bench_default.json')); print({k: d[k] for k in ('metric','value','unit','n_gpus','steps','warmup','ms_per_step','gpu_launches','vs_baseline','dtype')}); print(d['e2e']); print(d['roofline']); print(d['cpu_baseline']); print(d['clocks']); print(d['stages'])"
SECONDS=0; timeout -s KILL 900 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "reference wall ${SECONDS}s"; tail -1 gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json | cut -c1-600
