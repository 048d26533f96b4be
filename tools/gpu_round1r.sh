#!/bin/bash
mkdir -p gpurun_out
PROF_LLM_LAYERS=32 timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:decode_mega -s 1 -c 1 -o gpurun_out/prof_mega_r1 -f python tools/profile_step.py > gpurun_out/prof_mega.log 2>&1; echo "== ncu full exit $?"; tail -n 2 gpurun_out/prof_mega.log
