#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "llm_tiny or decode_batch or 7b_width" -p no:cacheprovider > gpurun_out/t_mega.log 2>&1; echo "== mega tests exit $?"; tail -n 5 gpurun_out/t_mega.log | cut -c1-400
for mk in mega legacy mega legacy; do
if [ $mk = legacy ]; then export VCL_NO_MEGAKERNEL=1; else unset VCL_NO_MEGAKERNEL; fi
timeout -s KILL 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_$mk.json 2> gpurun_out/bench_$mk.err; python -c "
import json; d=json.load(open('gpurun_out/bench_$mk.json')); s=d['stages']; print('$mk', round(d['value'],3), round(s['clip_ms'],2), round(s['prefill_ms'],2), round(s['decode_ms'],2), round(d['roofline']['frac'],3), d['gpu_launches'])"
done
