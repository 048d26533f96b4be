#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py tests/test_dropin_gpu.py -m gpu -q -x -k "gemv or llm or decode or 7b_width or config1 or generate or dropin" -p no:cacheprovider > gpurun_out/t_dec.log 2>&1; echo "== decode tests exit $?"; tail -n 4 gpurun_out/t_dec.log | cut -c1-400
VCL_DECODE_FUSED=1 timeout -s KILL 120 python tools/tc_trace.py 2>&1 | tail -9
run() { # name
timeout -s KILL 200 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_$1.json 2> gpurun_out/bench_$1.err; python -c "
import json; d=json.load(open('gpurun_out/bench_$1.json')); s=d['stages']; print('$1', round(d['value'],3), round(s['clip_ms'],2), round(s['prefill_ms'],2), round(s['decode_ms'],2), round(d['roofline']['frac'],3), d['gpu_launches'])" || tail -3 gpurun_out/bench_$1.err
}
VCL_DECODE_FUSED=1 run fused
run unfused
