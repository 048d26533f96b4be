#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 200 python tools/tc_trace.py 2>&1 | tail -9
