"""ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain-PyTorch restatement of the reference's video-conversation hot path, written as pure
functions over an HF/reference-style state_dict so the same weights can feed (a) this oracle,
(b) the real reference + HuggingFace modules (tests/golden/make_golden.py, run in the build
container where /root/reference exists) and (c) libvcl.so. Only tests/, __graft_entry__.smoke()
and bench.py's CPU-baseline / --impl reference legs may import this file; the product path
(video-llava_b200/) must never import it.

Parity status: the reference ships no tests or golden vectors, so this
restatement is pinned against outputs of the reference ITSELF: tests/golden/make_golden.py imports
/root/reference/video_chatgpt (+ the installed transformers 5.5.0 the reference delegates to, with
attn_implementation="eager") and commits small fixtures under tests/golden/; tests/test_oracle_cpu.py
checks this file against them. The pinned transformers@cae78c46 / torch 2.1 of the reference's
requirements.txt are not installable offline; known rounding differences are listed in SURVEY.md 8c.

Every function runs in the dtype of the tensors it is given (fp32 = gold, bf16 = "the reference's
own PyTorch path"), using one torch op per reference op so that bf16 rounding points coincide.

What each function follows (reference file:line; $TF = installed transformers):
  clip_hidden_states      $TF/models/clip/modeling_clip.py:138-218 (embeddings), 261-279 (eager
                          attention), 339-351 (MLP), 354-385 (layer), 667-693 (pre_layrnorm, encoder);
                          quick_gelu $TF/activations.py:117-123
  st_pool_torch           video_chatgpt/inference.py:13-44
  st_pool_numpy           scripts/save_spatio_temporal_clip_features.py:46-57
  project                 video_chatgpt/model/video_chatgpt.py:51-55,105; model/multimodal_projector/builder.py:33-51
  splice_embeddings       video_chatgpt/model/video_chatgpt.py:100-168 (use_vid_start_end branch :119-146)
  llm_forward             $TF/models/llama/modeling_llama.py:53-67 (RMSNorm), 124-168 (RoPE),
                          171-184 (MLP), 199-222 (eager attention), 292-331 (layer), 375-425 (model);
                          lm_head video_chatgpt/model/video_chatgpt.py:225-226
  greedy_generate         the hand-rolled greedy loop of SURVEY.md 9.2 (model.generate itself is broken
                          under transformers 5.x: video_chatgpt.py:253-257)
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch
import torch.nn.functional as F

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


@dataclass
class ClipCfg:
    hidden: int = 1024
    inter: int = 4096
    heads: int = 16
    layers: int = 24          # num_hidden_layers of the checkpoint (the path runs layers-1 of them)
    image: int = 224
    patch: int = 14
    eps: float = 1e-5


@dataclass
class LlmCfg:
    hidden: int = 4096
    inter: int = 11008
    heads: int = 32
    layers: int = 32
    vocab: int = 32003
    rms_eps: float = 1e-5
    rope_theta: float = 10000.0
    proj_type: str = "linear"  # or "mlp2x_gelu"
    mm_hidden: int = 1024
    vid_patch_token: int = 32000
    vid_start_token: int = 32001
    vid_end_token: int = 32002


# ---------------------------------------------------------------------------------------------
# frames -> pixel_values   (CLIPImageProcessor on 224x224 inputs reduces to this closed form)
# ---------------------------------------------------------------------------------------------
def preprocess_frames(frames_u8: np.ndarray | torch.Tensor) -> torch.Tensor:
    """[T,H,W,3] uint8 -> [T,3,H,W] fp32, (x/255 - mean)/std. video_chatgpt/inference.py:86."""
    x = torch.as_tensor(frames_u8).to(torch.float32) * (1.0 / 255.0)
    mean = torch.tensor(CLIP_MEAN, dtype=torch.float32)
    std = torch.tensor(CLIP_STD, dtype=torch.float32)
    x = (x - mean) / std
    return x.permute(0, 3, 1, 2).contiguous()


# ---------------------------------------------------------------------------------------------
# CLIP ViT
# ---------------------------------------------------------------------------------------------
def clip_hidden_states(sd: dict, cfg: ClipCfg, pixel_values: torch.Tensor, n_layers: int | None = None,
                       attn: str = "eager"):
    """Returns [hidden_states[0], ..., hidden_states[n_layers]] exactly as HF's
    CLIPVisionModel(..., output_hidden_states=True): [0] is the post-pre_layrnorm embedding,
    [i] the output of encoder layer i. The path consumes index cfg.layers-1 (hidden_states[-2])."""
    p = "vision_model."
    if n_layers is None:
        n_layers = cfg.layers - 1
    dt = sd[p + "embeddings.patch_embedding.weight"].dtype
    x = pixel_values.to(dt)
    n = x.shape[0]
    patch = F.conv2d(x, sd[p + "embeddings.patch_embedding.weight"], None, stride=cfg.patch)
    patch = patch.flatten(2).transpose(1, 2)                       # [N, P, C]
    cls = sd[p + "embeddings.class_embedding"].expand(n, 1, -1)
    h = torch.cat([cls, patch], dim=1) + sd[p + "embeddings.position_embedding.weight"][None]
    h = F.layer_norm(h, (cfg.hidden,), sd[p + "pre_layrnorm.weight"], sd[p + "pre_layrnorm.bias"], cfg.eps)
    out = [h]
    hd = cfg.hidden // cfg.heads
    scale = hd ** -0.5
    for l in range(n_layers):
        lp = f"{p}encoder.layers.{l}."
        r = h
        y = F.layer_norm(h, (cfg.hidden,), sd[lp + "layer_norm1.weight"], sd[lp + "layer_norm1.bias"], cfg.eps)
        q = F.linear(y, sd[lp + "self_attn.q_proj.weight"], sd[lp + "self_attn.q_proj.bias"])
        k = F.linear(y, sd[lp + "self_attn.k_proj.weight"], sd[lp + "self_attn.k_proj.bias"])
        v = F.linear(y, sd[lp + "self_attn.v_proj.weight"], sd[lp + "self_attn.v_proj.bias"])
        s = h.shape[1]
        q = q.view(n, s, cfg.heads, hd).transpose(1, 2)
        k = k.view(n, s, cfg.heads, hd).transpose(1, 2)
        v = v.view(n, s, cfg.heads, hd).transpose(1, 2)
        if attn == "sdpa":      # library-baseline timing only (bench.py); parity is defined on the eager form
            a = F.scaled_dot_product_attention(q, k, v, scale=scale).transpose(1, 2).reshape(n, s, cfg.hidden)
        else:
            w = torch.matmul(q, k.transpose(-1, -2)) * scale
            w = F.softmax(w, dim=-1, dtype=torch.float32).to(q.dtype)
            a = torch.matmul(w, v).transpose(1, 2).reshape(n, s, cfg.hidden)
        a = F.linear(a, sd[lp + "self_attn.out_proj.weight"], sd[lp + "self_attn.out_proj.bias"])
        h = r + a
        r = h
        y = F.layer_norm(h, (cfg.hidden,), sd[lp + "layer_norm2.weight"], sd[lp + "layer_norm2.bias"], cfg.eps)
        y = F.linear(y, sd[lp + "mlp.fc1.weight"], sd[lp + "mlp.fc1.bias"])
        y = y * torch.sigmoid(1.702 * y)                            # quick_gelu
        y = F.linear(y, sd[lp + "mlp.fc2.weight"], sd[lp + "mlp.fc2.bias"])
        h = r + y
        out.append(h)
    return out


# ---------------------------------------------------------------------------------------------
# spatio-temporal pooling
# ---------------------------------------------------------------------------------------------
def st_pool_torch(features: torch.Tensor) -> torch.Tensor:
    """[T,P,C] -> [100+P, C] fp16. Mirrors the op sequence of inference.py:13-44 (including the fp32
    zero padding that promotes the temporal half when T < 100, and the final .half())."""
    t, s, c = features.shape
    temporal = torch.mean(features, dim=1)
    pad = 100 - t
    if pad > 0:
        temporal = torch.cat((temporal, torch.zeros(pad, c, device=features.device)), dim=0)
    spatial = torch.mean(features, dim=0)
    return torch.cat([temporal, spatial], dim=0).half()


def st_pool_numpy(features: np.ndarray, num_temporal_tokens: int = 100) -> np.ndarray:
    """numpy twin used by the offline feature extractor (scripts/...:46-57)."""
    t, s, c = features.shape
    temporal = np.mean(features, axis=1)
    pad = num_temporal_tokens - t
    if pad > 0:
        temporal = np.pad(temporal, ((0, pad), (0, 0)), mode="constant")
    spatial = np.mean(features, axis=0)
    return np.concatenate([temporal, spatial], axis=0)


# ---------------------------------------------------------------------------------------------
# projector + splice + LLaMA
# ---------------------------------------------------------------------------------------------
def project(sd: dict, cfg: LlmCfg, feats: torch.Tensor) -> torch.Tensor:
    if cfg.proj_type == "linear":
        return F.linear(feats, sd["model.mm_projector.weight"], sd["model.mm_projector.bias"])
    y = F.linear(feats, sd["model.mm_projector.0.weight"], sd["model.mm_projector.0.bias"])
    y = F.gelu(y)
    return F.linear(y, sd["model.mm_projector.2.weight"], sd["model.mm_projector.2.bias"])


def splice_embeddings(sd: dict, cfg: LlmCfg, ids: torch.Tensor, feats: torch.Tensor | None) -> torch.Tensor:
    """embed_tokens(ids) with the projected video rows replacing the <vid_patch> span that follows
    <vid_start> (whose own embedding, like <vid_end>'s, is kept). Raises ValueError like the reference
    on malformed spans."""
    emb = F.embedding(ids, sd["model.embed_tokens.weight"])
    if feats is None:
        return emb
    vid = project(sd, cfg, feats.to(emb.dtype))
    rows = []
    for b in range(ids.shape[0]):
        cur = ids[b]
        if (cur == cfg.vid_patch_token).sum() == 0:
            rows.append(emb[b])
            continue
        if (cur == cfg.vid_start_token).sum() != (cur == cfg.vid_end_token).sum():
            raise ValueError("The number of video start tokens and video end tokens should be the same.")
        starts = torch.where(cur == cfg.vid_start_token)[0]
        e = emb[b]
        for s in starts:
            s = int(s)
            n = vid.shape[1]
            if cur[s + n + 1] != cfg.vid_end_token:
                raise ValueError("The video end token should follow the video start token.")
            e = torch.cat((e[: s + 1], vid[b], e[s + n + 1:]), dim=0)
        rows.append(e)
    return torch.stack(rows, dim=0)


def _rmsnorm(x, w, eps):
    dt = x.dtype
    xf = x.to(torch.float32)
    xf = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)
    return w * xf.to(dt)


def _rope_cos_sin(cfg: LlmCfg, positions: torch.Tensor, dtype, device):
    hd = cfg.hidden // cfg.heads
    inv = 1.0 / (cfg.rope_theta ** (torch.arange(0, hd, 2, dtype=torch.int64).to(dtype=torch.float) / hd))
    freqs = positions.to(torch.float32)[:, None] * inv[None, :]
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos().to(dtype=dtype, device=device), emb.sin().to(dtype=dtype, device=device)


def _rotate_half(x):
    h = x.shape[-1] // 2
    return torch.cat((-x[..., h:], x[..., :h]), dim=-1)


def llm_forward(sd: dict, cfg: LlmCfg, ids: torch.Tensor, feats: torch.Tensor | None = None,
                past: list | None = None, n_layers: int | None = None, all_logits: bool = False,
                attn: str = "eager"):
    """One forward of VideoChatGPTLlamaForCausalLM. `past` is a list of (k, v) per layer
    ([B,H,S,hd]); with a past the video features are ignored exactly as the reference does for
    single-token inputs (video_chatgpt.py:103). Returns (logits, hidden_states, past) where
    hidden_states[i] is HF's hidden_states[i] for i < L and the post-norm output for i = L, and
    logits covers the last position only unless all_logits."""
    if n_layers is None:
        n_layers = cfg.layers
    b, s = ids.shape
    if ids.shape[1] != 1 and feats is not None:
        h = splice_embeddings(sd, cfg, ids, feats)
    else:
        h = F.embedding(ids, sd["model.embed_tokens.weight"])
    dt = h.dtype
    p0 = 0 if past is None else past[0][0].shape[2]
    pos = torch.arange(p0, p0 + s)
    cos, sin = _rope_cos_sin(cfg, pos, dt, h.device)
    cos, sin = cos[None, None], sin[None, None]
    hd = cfg.hidden // cfg.heads
    scale = hd ** -0.5
    mask = None
    if s > 1:
        full = torch.full((s, p0 + s), torch.finfo(dt).min, dtype=dt, device=h.device)
        mask = torch.triu(full, diagonal=p0 + 1)
    hs = [h]
    new_past = []
    for l in range(n_layers):
        lp = f"model.layers.{l}."
        r = h
        y = _rmsnorm(h, sd[lp + "input_layernorm.weight"], cfg.rms_eps)
        q = F.linear(y, sd[lp + "self_attn.q_proj.weight"]).view(b, s, cfg.heads, hd).transpose(1, 2)
        k = F.linear(y, sd[lp + "self_attn.k_proj.weight"]).view(b, s, cfg.heads, hd).transpose(1, 2)
        v = F.linear(y, sd[lp + "self_attn.v_proj.weight"]).view(b, s, cfg.heads, hd).transpose(1, 2)
        q = (q * cos) + (_rotate_half(q) * sin)
        k = (k * cos) + (_rotate_half(k) * sin)
        if past is not None:
            k = torch.cat([past[l][0], k], dim=2)
            v = torch.cat([past[l][1], v], dim=2)
        new_past.append((k, v))
        if attn == "sdpa":
            # what transformers 5.x picks by default ($TF/integrations/sdpa_attention.py); used only to
            # screen prompts for eager-vs-sdpa self-agreement, never as the parity target
            a = F.scaled_dot_product_attention(q, k, v, attn_mask=mask, scale=scale)
            a = a.transpose(1, 2).reshape(b, s, cfg.hidden)
        else:
            w = torch.matmul(q, k.transpose(2, 3)) * scale
            if mask is not None:
                w = w + mask
            w = F.softmax(w, dim=-1, dtype=torch.float32).to(q.dtype)
            a = torch.matmul(w, v).transpose(1, 2).reshape(b, s, cfg.hidden)
        h = r + F.linear(a, sd[lp + "self_attn.o_proj.weight"])
        r = h
        y = _rmsnorm(h, sd[lp + "post_attention_layernorm.weight"], cfg.rms_eps)
        y = F.silu(F.linear(y, sd[lp + "mlp.gate_proj.weight"])) * F.linear(y, sd[lp + "mlp.up_proj.weight"])
        h = r + F.linear(y, sd[lp + "mlp.down_proj.weight"])
        hs.append(h)
    logits = None
    if n_layers == cfg.layers:
        hn = _rmsnorm(h, sd["model.norm.weight"], cfg.rms_eps)
        hs[-1] = hn
        logits = F.linear(hn if all_logits else hn[:, -1:], sd["lm_head.weight"])
    return logits, hs, new_past


def greedy_generate(sd: dict, cfg: LlmCfg, ids: torch.Tensor, feats: torch.Tensor, n_new: int,
                    forced: torch.Tensor | None = None, attn: str = "eager"):
    """Greedy decoding with a KV cache, EOS ignored. If `forced` ([B, n_new]) is given the loop is
    teacher-forced with those tokens (the arg-max of every step is still reported).
    Returns (tokens [B,n_new] int64, last-position logits per step [n_new,B,V] fp32)."""
    toks, logs = [], []
    logits, _, past = llm_forward(sd, cfg, ids, feats, attn=attn)
    for i in range(n_new):
        lg = logits[:, -1].float()
        logs.append(lg)
        nxt = lg.argmax(-1)
        toks.append(nxt)
        if i + 1 == n_new:
            break
        feed = nxt if forced is None else forced[:, i].to(nxt.device)
        logits, _, past = llm_forward(sd, cfg, feed[:, None], feats, past, attn=attn)
    return torch.stack(toks, 1), torch.stack(logs, 0)


# ---------------------------------------------------------------------------------------------
# synthetic inputs / weights shared by tests, the golden generator and bench.py
# ---------------------------------------------------------------------------------------------
def make_prompt_ids(cfg: LlmCfg, n_vid: int, seed: int = 1, n_pre: int = 63, n_post: int = 26,
                    batch: int = 1) -> torch.Tensor:
    """[1] + n_pre random ids + <vid_start> + <vid_patch>*n_vid + <vid_end> + n_post random ids
    (1+63+1+356+1+26 = 448)."""
    g = torch.Generator().manual_seed(seed)
    rows = []
    for _ in range(batch):
        pre = torch.randint(3, 32000 if cfg.vocab > 32000 else cfg.vocab - 3, (n_pre,), generator=g)
        post = torch.randint(3, 32000 if cfg.vocab > 32000 else cfg.vocab - 3, (n_post,), generator=g)
        rows.append(torch.cat([torch.tensor([1]), pre, torch.tensor([cfg.vid_start_token]),
                               torch.full((n_vid,), cfg.vid_patch_token), torch.tensor([cfg.vid_end_token]),
                               post]))
    return torch.stack(rows, 0).to(torch.int64)


def make_frames(clip: int, t: int, size: int = 224) -> np.ndarray:
    return np.random.default_rng(1000 + clip).integers(0, 256, (t, size, size, 3), dtype=np.uint8)


def random_clip_state(cfg: ClipCfg, seed: int = 0, dtype=torch.float32, n_layers: int | None = None) -> dict:
    g = torch.Generator().manual_seed(seed)
    r = lambda *s, std=0.02: (torch.randn(*s, generator=g) * std)
    p = "vision_model."
    P = (cfg.image // cfg.patch) ** 2
    sd = {
        p + "embeddings.class_embedding": r(cfg.hidden, std=cfg.hidden ** -0.5),
        p + "embeddings.patch_embedding.weight": r(cfg.hidden, 3, cfg.patch, cfg.patch, std=0.02),
        p + "embeddings.position_embedding.weight": r(P + 1, cfg.hidden, std=0.02),
        p + "pre_layrnorm.weight": 1 + r(cfg.hidden, std=0.05),
        p + "pre_layrnorm.bias": r(cfg.hidden, std=0.02),
    }
    nl = cfg.layers if n_layers is None else n_layers
    for l in range(nl):
        lp = f"{p}encoder.layers.{l}."
        for nm in ("q_proj", "k_proj", "v_proj", "out_proj"):
            sd[lp + f"self_attn.{nm}.weight"] = r(cfg.hidden, cfg.hidden, std=cfg.hidden ** -0.5)
            sd[lp + f"self_attn.{nm}.bias"] = r(cfg.hidden, std=0.02)
        sd[lp + "mlp.fc1.weight"] = r(cfg.inter, cfg.hidden, std=cfg.hidden ** -0.5)
        sd[lp + "mlp.fc1.bias"] = r(cfg.inter, std=0.02)
        sd[lp + "mlp.fc2.weight"] = r(cfg.hidden, cfg.inter, std=cfg.inter ** -0.5)
        sd[lp + "mlp.fc2.bias"] = r(cfg.hidden, std=0.02)
        for nm in ("layer_norm1", "layer_norm2"):
            sd[lp + nm + ".weight"] = 1 + r(cfg.hidden, std=0.05)
            sd[lp + nm + ".bias"] = r(cfg.hidden, std=0.02)
    return {k: v.to(dtype) for k, v in sd.items()}


def random_llm_state(cfg: LlmCfg, seed: int = 0, dtype=torch.float32) -> dict:
    g = torch.Generator().manual_seed(seed)
    r = lambda *s, std=0.02: (torch.randn(*s, generator=g) * std)
    d, f = cfg.hidden, cfg.inter
    sd = {
        "model.embed_tokens.weight": r(cfg.vocab, d, std=1.0),
        "model.norm.weight": 1 + r(d, std=0.05),
        "lm_head.weight": r(cfg.vocab, d, std=d ** -0.5),
    }
    if cfg.proj_type == "linear":
        sd["model.mm_projector.weight"] = r(d, cfg.mm_hidden, std=cfg.mm_hidden ** -0.5)
        sd["model.mm_projector.bias"] = r(d, std=0.02)
    else:
        sd["model.mm_projector.0.weight"] = r(d, cfg.mm_hidden, std=cfg.mm_hidden ** -0.5)
        sd["model.mm_projector.0.bias"] = r(d, std=0.02)
        sd["model.mm_projector.2.weight"] = r(d, d, std=d ** -0.5)
        sd["model.mm_projector.2.bias"] = r(d, std=0.02)
    for l in range(cfg.layers):
        lp = f"model.layers.{l}."
        for nm in ("q_proj", "k_proj", "v_proj", "o_proj"):
            sd[lp + f"self_attn.{nm}.weight"] = r(d, d, std=d ** -0.5)
        sd[lp + "mlp.gate_proj.weight"] = r(f, d, std=d ** -0.5)
        sd[lp + "mlp.up_proj.weight"] = r(f, d, std=d ** -0.5)
        sd[lp + "mlp.down_proj.weight"] = r(d, f, std=f ** -0.5)
        sd[lp + "input_layernorm.weight"] = 1 + r(d, std=0.05)
        sd[lp + "post_attention_layernorm.weight"] = 1 + r(d, std=0.05)
    return {k: v.to(dtype) for k, v in sd.items()}


def device_llm_state(cfg: LlmCfg, device, seed: int = 0, dtype=torch.bfloat16) -> dict:
    """Full-size random-init LLM weights generated ON the device in `dtype` (a 7B fp32 state built on
    the host takes minutes and 27 GB); same distributions as random_llm_state, different bytes."""
    g = torch.Generator(device=device).manual_seed(seed)
    big = lambda r, c, std: torch.empty(r, c, device=device, dtype=dtype).normal_(0.0, std, generator=g)
    vec = lambda n, std: torch.empty(n, device=device, dtype=torch.float32).normal_(0.0, std, generator=g)
    d, f = cfg.hidden, cfg.inter
    sd = {"model.embed_tokens.weight": big(cfg.vocab, d, 1.0), "model.norm.weight": (1 + vec(d, 0.05)).to(dtype),
          "lm_head.weight": big(cfg.vocab, d, d ** -0.5)}
    if cfg.proj_type == "linear":
        sd["model.mm_projector.weight"] = big(d, cfg.mm_hidden, cfg.mm_hidden ** -0.5)
        sd["model.mm_projector.bias"] = vec(d, 0.02).to(dtype)
    else:
        sd["model.mm_projector.0.weight"] = big(d, cfg.mm_hidden, cfg.mm_hidden ** -0.5)
        sd["model.mm_projector.0.bias"] = vec(d, 0.02).to(dtype)
        sd["model.mm_projector.2.weight"] = big(d, d, d ** -0.5)
        sd["model.mm_projector.2.bias"] = vec(d, 0.02).to(dtype)
    for l in range(cfg.layers):
        lp = f"model.layers.{l}."
        for nm in ("q_proj", "k_proj", "v_proj", "o_proj"):
            sd[lp + f"self_attn.{nm}.weight"] = big(d, d, d ** -0.5)
        sd[lp + "mlp.gate_proj.weight"] = big(f, d, d ** -0.5)
        sd[lp + "mlp.up_proj.weight"] = big(f, d, d ** -0.5)
        sd[lp + "mlp.down_proj.weight"] = big(d, f, f ** -0.5)
        sd[lp + "input_layernorm.weight"] = (1 + vec(d, 0.05)).to(dtype)
        sd[lp + "post_attention_layernorm.weight"] = (1 + vec(d, 0.05)).to(dtype)
    return sd
