#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python tools/mega_trace.py 2>&1 | tail -12
