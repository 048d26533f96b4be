#!/bin/bash
mkdir -p gpurun_out
PROF_LLM_LAYERS=2 timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1e.csv python tools/profile_step.py > gpurun_out/prof_step.log 2>&1; echo "== ncu exit $?"; tail -n 2 gpurun_out/prof_step.log; wc -l gpurun_out/launches_r1e.csv
