# scratch: the GPU session currently queued (tools/gpurun_retry.sh --timeout 2400 -- 'bash tools/_session.sh')
bash tools/gpu_run.sh "tests:attention or prefill or llm_tiny or two_layers"
bash tools/gpu_run.sh ab:VCL_PREFILL_ATTN_FLASH=1 ab:X=1 ab:VCL_PREFILL_ATTN_FLASH=1 ab:X=2
PROF_SKIP_VIT=1 NCU_SKIP=0 NCU_COUNT=12 bash tools/gpu_run.sh 'ncufull:gemm_bf16|attn_fwd|attn_prefill|rope_kv|rownorm_warp|embed_splice'
AB_ARGS="--config 3" AB_STEPS=3 bash tools/gpu_run.sh ab:VCL_PREFILL_ATTN_FLASH=1 ab:X=3
