#!/bin/bash
mkdir -p gpurun_out
PROF_LLM_LAYERS=2 PROF_CLIPS=16 timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:gemv_mma -s 10 -c 4 -o gpurun_out/prof_gmma_r1 -f python tools/profile_step.py > gpurun_out/prof_gmma.log 2>&1; echo "== ncu full exit $?"; tail -n 2 gpurun_out/prof_gmma.log
