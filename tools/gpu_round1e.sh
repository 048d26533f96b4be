#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "== smoke exit $?"; tail -n 5 gpurun_out/smoke.log
timeout -s KILL 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench1.json 2> gpurun_out/bench1.err; echo "== bench exit $?"; tail -n 5 gpurun_out/bench1.err; cat gpurun_out/bench1.json
