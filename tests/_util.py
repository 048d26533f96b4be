"""Helpers shared by the GPU parity tests: build a vcl Engine from oracle configs."""
import torch

import vcl_native as vn
from oracle import vcl_oracle as O


def make_engine(clip: O.ClipCfg | None = None, llm: O.LlmCfg | None = None, clip_run_layers=None,
                max_frames=1, max_batch=1, max_seq=512):
    clip = clip or O.ClipCfg()
    llm = llm or O.LlmCfg(hidden=512, inter=1024, heads=4, layers=0)
    c = vn.vcl_config()
    c.clip_layers = (clip.layers - 1) if clip_run_layers is None else clip_run_layers
    c.clip_hidden, c.clip_inter, c.clip_heads = clip.hidden, clip.inter, clip.heads
    c.image_size, c.patch_size, c.clip_ln_eps = clip.image, clip.patch, clip.eps
    c.llm_layers, c.llm_hidden, c.llm_inter, c.llm_heads = llm.layers, llm.hidden, llm.inter, llm.heads
    c.vocab, c.rms_eps, c.rope_theta = llm.vocab, llm.rms_eps, llm.rope_theta
    c.proj_type = vn.PROJ_LINEAR if llm.proj_type == "linear" else vn.PROJ_MLP2X_GELU
    c.n_temporal = 100
    c.max_frames, c.max_batch, c.max_seq = max_frames, max_batch, max_seq
    return vn.Engine(c)


def to_dev(sd, dtype=torch.bfloat16, device="cuda"):
    return {k: v.to(device=device, dtype=dtype) for k, v in sd.items()}


def relerr(a, b):
    a = a.float()
    b = b.float().to(a.device)
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def vid_start_of(ids, llm: O.LlmCfg):
    """index of <vid_start> per row (-1 if absent), int32 on the ids' device"""
    out = []
    for row in ids.tolist():
        out.append(row.index(llm.vid_start_token) if llm.vid_start_token in row else -1)
    return torch.tensor(out, dtype=torch.int32, device=ids.device)
