# scratch: the GPU session currently queued (tools/gpurun_retry.sh --timeout 2400 -- 'bash tools/_session.sh')
bash tools/gpu_run.sh tests
bash tools/gpu_run.sh ab:VCL_PREFILL_ROPE_SEPARATE=1 ab:X=1 ab:VCL_PREFILL_ROPE_SEPARATE=1 ab:X=2
AB_ARGS="--config 3" AB_STEPS=3 bash tools/gpu_run.sh ab:VCL_PREFILL_ROPE_SEPARATE=1 ab:X=3
