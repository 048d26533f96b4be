#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py -m gpu -q -k "attention_vit or clip or config1" -p no:cacheprovider > gpurun_out/t_k.log 2>&1; echo "== tests exit $?"; tail -n 8 gpurun_out/t_k.log | cut -c1-400
timeout -s KILL 300 python tools/microbench.py attn 2>&1 | tail -1
VCL_ATTN_TWO_TILE=1 timeout -s KILL 300 python tools/microbench.py attn 2>&1 | tail -1
for v in new old new old; do
if [ $v = old ]; then export VCL_ATTN_TWO_TILE=1; else unset VCL_ATTN_TWO_TILE; fi
timeout -s KILL 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_a$v.json 2> gpurun_out/bench_a$v.err; python -c "
import json; d=json.load(open('gpurun_out/bench_a$v.json')); s=d['stages']; print('$v', round(d['value'],3), round(s['clip_ms'],2), round(s['prefill_ms'],2), round(s['decode_ms'],2))"
done
