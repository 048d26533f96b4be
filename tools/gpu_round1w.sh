#!/bin/bash
mkdir -p gpurun_out
PROF_LLM_LAYERS=2 timeout -s KILL 1200 ncu --set full --clock-control none --import-source on -k "regex:gemm_bf16_tn|attn_vit_tc|st_pool|decode_attn_cluster|rownorm_warp|gemv_kernel" -s 170 -c 60 -o gpurun_out/prof_mix_r1 -f python tools/profile_step.py > gpurun_out/prof_mix.log 2>&1; echo "== ncu full exit $?"; tail -n 2 gpurun_out/prof_mix.log; ls -la gpurun_out/prof_mix_r1.ncu-rep
