#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 200 python tools/tc_trace.py 2>&1 | tail -9
timeout -s KILL 300 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py -m gpu -q -k "gemv or llm or decode or 7b_width or config1" -p no:cacheprovider > gpurun_out/t_gemv.log 2>&1; echo "== gemv tests exit $?"; tail -n 3 gpurun_out/t_gemv.log | cut -c1-400
run() { # name
timeout -s KILL 300 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_$1.json 2> gpurun_out/bench_$1.err; python -c "
import json; d=json.load(open('gpurun_out/bench_$1.json')); s=d['stages']; print('$1', round(d['value'],3), round(s['clip_ms'],2), round(s['prefill_ms'],2), round(s['decode_ms'],2), round(d['roofline']['frac'],3), d['gpu_launches'])" || tail -3 gpurun_out/bench_$1.err
}
run tc113
VCL_GEMV_TC_SMEM_KB=150 run tc150
VCL_GEMV_LEGACY=1 run legacy
