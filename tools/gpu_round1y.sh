#!/bin/bash
mkdir -p gpurun_out
VCL_MEGAKERNEL=1 timeout -s KILL 300 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "llm_tiny or decode_batch or 7b_width" -p no:cacheprovider > gpurun_out/t_mega.log 2>&1; echo "== mega tests exit $?"; tail -n 5 gpurun_out/t_mega.log | cut -c1-400
timeout -s KILL 300 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py -m gpu -q -k "attention_vit or clip or config1" -p no:cacheprovider > gpurun_out/t_k.log 2>&1; echo "== attn tests exit $?"; tail -n 8 gpurun_out/t_k.log | cut -c1-400
timeout -s KILL 300 python tools/microbench.py attn 2>&1 | tail -1
VCL_ATTN_TWO_TILE=1 timeout -s KILL 300 python tools/microbench.py attn 2>&1 | tail -1
for v in new_legacy old_legacy new_mega new_legacy old_legacy new_mega; do
unset VCL_ATTN_TWO_TILE VCL_MEGAKERNEL
case $v in old_*) export VCL_ATTN_TWO_TILE=1;; esac
case $v in *_mega) export VCL_MEGAKERNEL=1;; esac
timeout -s KILL 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err; python -c "
import json; d=json.load(open('gpurun_out/bench_$v.json')); s=d['stages']; print('$v', round(d['value'],3), round(s['clip_ms'],2), round(s['prefill_ms'],2), round(s['decode_ms'],2), round(d['roofline']['frac'],3), d['gpu_launches'])" || tail -3 gpurun_out/bench_$v.err
done
