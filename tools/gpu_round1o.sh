#!/bin/bash
mkdir -p gpurun_out
PROF_LLM_LAYERS=2 timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1d.csv python tools/profile_step.py > gpurun_out/prof_step.log 2>&1; echo "== ncu exit $?"; tail -n 2 gpurun_out/prof_step.log
for per in 1 2 1 2; do
VCL_GEMV_CTAS_PER_SM=$per timeout -s KILL 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_p$per.json 2> gpurun_out/bench_p$per.err; python -c "
import json; d=json.load(open('gpurun_out/bench_p$per.json')); s=d['stages']; print('per_sm=$per', round(d['value'],3), round(s['clip_ms'],2), round(s['prefill_ms'],2), round(s['decode_ms'],2), round(d['roofline']['frac'],3))"
done
