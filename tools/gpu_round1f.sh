#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_dropin_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/t_dropin.log 2>&1; echo "== dropin exit $?"; tail -n 15 gpurun_out/t_dropin.log
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1a.csv python tools/profile_step.py > gpurun_out/prof_step.log 2>&1; echo "== ncu exit $?"; tail -n 3 gpurun_out/prof_step.log; wc -l gpurun_out/launches_r1a.csv
