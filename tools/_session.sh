# scratch: the GPU session currently queued (tools/gpurun_retry.sh --timeout 2400 -- 'bash tools/_session.sh')
bash tools/gpu_run.sh tests smoke bench ref launches
NCU_COUNT=12 NCU_SKIP=38 bash tools/gpu_run.sh 'ncufull:gemv_tc_kernel|decode_attn|argmax' sass
PROF_SKIP_VIT=1 NCU_SKIP=0 NCU_COUNT=11 bash tools/gpu_run.sh 'ncufull:gemm_bf16|attn_fwd|attn_prefill|rope_kv|rownorm_warp|embed_splice'
bash tools/gpu_run.sh bench:--config,3,--no-cpu,--no-library bench:--config,4,--no-cpu,--no-library
