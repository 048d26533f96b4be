"""One eager pass of the hot path at 7B width (few LLM layers) for ncu launch lists:
ViT over 100 frames -> pool -> prefill -> 2 decode steps (no CUDA graph so that every kernel is listed)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "video-llava_b200"))
import bench  # noqa: E402
import vcl_native as vn  # noqa: E402

L = int(os.environ.get("PROF_LLM_LAYERS", "4"))
B = int(os.environ.get("PROF_CLIPS", "1"))
bench.MODELS["7b"]["layers"] = L
dev = torch.device("cuda:0")
c = vn.vcl_config()
c.clip_layers, c.clip_hidden, c.clip_inter, c.clip_heads = 23, 1024, 4096, 16
c.image_size, c.patch_size, c.clip_ln_eps = 224, 14, 1e-5
c.llm_layers, c.llm_hidden, c.llm_inter, c.llm_heads = L, 4096, 11008, 32
c.vocab, c.rms_eps, c.rope_theta = 32003, 1e-5, 10000.0
c.proj_type, c.n_temporal = vn.PROJ_LINEAR, 100
c.max_frames, c.max_batch, c.max_seq = 100, B, 480
eng = vn.Engine(c)
clip_sd, llm_sd = bench.device_weights("7b", dev)
eng.load_clip(clip_sd); eng.load_llm(llm_sd)
del clip_sd, llm_sd
frames = torch.randint(0, 256, (100, 224, 224, 3), dtype=torch.uint8, device=dev)
ids = torch.randint(3, 32000, (B, 448), device=dev); ids[:, 64] = 32001; ids[:, 65:421] = 32000; ids[:, 421] = 32002
vs = torch.full((B,), 64, dtype=torch.int32, device=dev)
feats = torch.empty(B, 356, 1024, dtype=torch.bfloat16, device=dev)
for it in range(2):
    for b in range(B if not os.environ.get("PROF_SKIP_VIT") else 0):
        eng.clip_features(frames, out=feats[b])
    _, _, tok = eng.prefill(ids, feats, vs)
    for i in range(2):
        _, tok = eng.decode_step(tok, 448 + i)
torch.cuda.synchronize()
print("done", tok.tolist())
