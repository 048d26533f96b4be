#!/bin/bash
mkdir -p gpurun_out
VCL_MEGAKERNEL=1 timeout -s KILL 200 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "llm_tiny or decode_batch or 7b_width" -p no:cacheprovider > gpurun_out/t_mega.log 2>&1; echo "== mega tests exit $?"; tail -n 3 gpurun_out/t_mega.log | cut -c1-400
VCL_GEMV_L2_ROWS=16 timeout -s KILL 200 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py -m gpu -q -x -k "gemv or llm_tiny or 7b_width" -p no:cacheprovider > gpurun_out/t_gemv.log 2>&1; echo "== gemv tests exit $?"; tail -n 3 gpurun_out/t_gemv.log | cut -c1-400
run() { # name
timeout -s KILL 300 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_$1.json 2> gpurun_out/bench_$1.err; python -c "
import json; d=json.load(open('gpurun_out/bench_$1.json')); s=d['stages']; print('$1', round(d['value'],3), round(s['clip_ms'],2), round(s['prefill_ms'],2), round(s['decode_ms'],2), round(d['roofline']['frac'],3), d['gpu_launches'])" || tail -3 gpurun_out/bench_$1.err
}
for r in 0 8 16 32 64; do export VCL_GEMV_L2_ROWS=$r; run legacy_l2r$r; done
unset VCL_GEMV_L2_ROWS
export VCL_MEGAKERNEL=1
for r in 0 12 24 48; do export VCL_MEGA_L2_SLOTS=$r; run mega_l2s$r; done
