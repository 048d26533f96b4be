#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -m gpu -q -x -p no:cacheprovider -k "gemv or decode or llm or dropin or 7b or variants" > gpurun_out/t_dec.log 2>&1; echo "== tests exit $?"; tail -n 5 gpurun_out/t_dec.log | cut -c1-400
run() { # name args
n=$1; shift
timeout -s KILL 400 python bench.py --steps 3 --warmup 3 --no-cpu "$@" > gpurun_out/bench_$n.json 2> gpurun_out/bench_$n.err; python -c "
import json; d=json.load(open('gpurun_out/bench_$n.json')); s=d['stages']; print('$n', round(d['value'],3), round(s['clip_ms'],2), round(s['prefill_ms'],2), round(s['decode_ms'],2), round(d['roofline']['frac'],3), d['gpu_launches'])" || tail -3 gpurun_out/bench_$n.err
}
run 13b_b4 --model 13b --clips 4
run 7b_b4 --model 7b --clips 4
run 7b_b2 --model 7b --clips 2
run 7b_b1 --model 7b --clips 1
