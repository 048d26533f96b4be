#!/bin/bash
mkdir -p gpurun_out
PROF_LLM_LAYERS=4 timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1b.csv python tools/profile_step.py > gpurun_out/prof_step.log 2>&1; echo "== ncu exit $?"; tail -n 2 gpurun_out/prof_step.log
PROF_LLM_LAYERS=4 timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:gemv_kernel -s 22 -c 6 -o gpurun_out/prof_gemv_r1 -f python tools/profile_step.py > gpurun_out/prof_gemv.log 2>&1; echo "== ncu full exit $?"; tail -n 2 gpurun_out/prof_gemv.log; ls -la gpurun_out/*.ncu-rep
