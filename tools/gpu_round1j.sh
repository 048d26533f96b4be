#!/bin/bash
mkdir -p gpurun_out
PROF_LLM_LAYERS=2 timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1c.csv python tools/profile_step.py > gpurun_out/prof_step.log 2>&1; echo "== ncu exit $?"; tail -n 2 gpurun_out/prof_step.log
PROF_LLM_LAYERS=2 timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:attn_vit_tc -s 2 -c 2 -o gpurun_out/prof_attn_r1 -f python tools/profile_step.py > gpurun_out/prof_attn.log 2>&1; echo "== ncu full exit $?"; tail -n 2 gpurun_out/prof_attn.log
