#!/bin/bash
mkdir -p gpurun_out
SECONDS=0
timeout -s KILL 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=6 > gpurun_out/t_all.log 2>&1; echo "== all gpu tests exit $? in ${SECONDS}s"; tail -n 12 gpurun_out/t_all.log | cut -c1-200
timeout -s KILL 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
SECONDS=0; timeout -s KILL 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench wall ${SECONDS}s"; python -c "
import json; d=json.load(open('gpurun_out/bench_default.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['stages']['clip_ms'], d['stages']['prefill_ms'], d['stages']['decode_ms'], d['clocks'])"
