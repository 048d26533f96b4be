"""The reference-facing Python surface (video_chatgpt.* mirror) on the GPU: same calls the
reference's callers make (inference.py:93-120, scripts/save_spatio_temporal_clip_features.py),
checked against the golden fixtures and the oracle."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import vcl_oracle as O  # noqa: E402

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _model(llm_cfg, clip_layers_total=3, max_batch=2):
    from video_chatgpt.model import VideoChatGPTConfig, VideoChatGPTLlamaForCausalLM
    cfg = VideoChatGPTConfig(hidden_size=llm_cfg.hidden, intermediate_size=llm_cfg.inter,
                             num_hidden_layers=llm_cfg.layers, num_attention_heads=llm_cfg.heads,
                             vocab_size=llm_cfg.vocab, use_mm_proj=True, mm_hidden_size=1024)
    clip = dict(hidden_size=1024, intermediate_size=1024, num_hidden_layers=clip_layers_total, num_attention_heads=16)
    m = VideoChatGPTLlamaForCausalLM(cfg, clip_config=clip, max_batch=max_batch, max_seq=480)
    vc = m.get_model().vision_config
    vc.vid_patch_token, vc.vid_start_token, vc.vid_end_token, vc.use_vid_start_end = 32000, 32001, 32002, True
    return m


def test_pool_functions_vs_reference_fixtures():
    from video_chatgpt.inference import get_spatio_temporal_features_torch
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "vcl_save_feats", os.path.join(os.path.dirname(G), "..", "video-llava_b200", "scripts",
                                       "save_spatio_temporal_clip_features.py"))
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    g = np.load(os.path.join(G, "pool.npz"))
    gen = torch.Generator().manual_seed(5)
    f8 = torch.randn(8, 256, 1024, generator=gen).half()
    f100 = torch.randn(100, 256, 1024, generator=gen).half()
    o8 = get_spatio_temporal_features_torch(f8.cuda())
    assert o8.dtype == torch.float16 and o8.shape == (356, 1024) and o8.is_cuda
    assert (o8[8:100] == 0).all()
    for ours, ref in [(o8.cpu().numpy(), g["t8_torch"]),
                      (mod.get_spatio_temporal_features(f8.numpy()), g["t8_numpy"]),
                      (get_spatio_temporal_features_torch(f100.cuda()).cpu().numpy()[::7], g["t100_torch_rows"]),
                      (get_spatio_temporal_features_torch(f100.bfloat16().cuda()).cpu().numpy()[::7], g["t100_bf16_rows"])]:
        assert ours.dtype == np.float16 and ours.shape == ref.shape
        d = np.abs(ours.astype(np.float32) - ref.astype(np.float32))
        ulp = np.maximum(np.abs(ref.astype(np.float32)), 2.0 ** -14) * 2.0 ** -10
        # fp32 summation order differs from torch's / numpy's trees: <= 1 ulp of the output, mostly exact
        assert (d <= ulp * (8 if ref is g["t100_bf16_rows"] else 1)).all(), d.max()
        assert (ours == ref).mean() > 0.97


@torch.no_grad()
def test_tower_pool_generate_like_the_reference_caller():
    from video_chatgpt.inference import get_spatio_temporal_features_torch
    from video_chatgpt.model.utils import KeywordsStoppingCriteria
    lcfg = O.LlmCfg(hidden=512, inter=1024, heads=4, layers=2)
    ccfg = O.ClipCfg(hidden=1024, inter=1024, heads=16, layers=3)
    m = _model(lcfg, 3, max_batch=1)
    lsd, csd = O.random_llm_state(lcfg, seed=21), O.random_clip_state(ccfg, seed=11)
    m.load_state_dict(lsd)
    tower = m.get_vision_tower()
    tower.load_state_dict(csd)
    px = O.preprocess_frames(O.make_frames(7, 3))
    outs = tower(px.half().cuda(), output_hidden_states=True)
    assert len(outs.hidden_states) == 4
    feats_in = outs.hidden_states[-2][:, 1:]
    assert feats_in.shape == (3, 256, 1024)
    with pytest.raises(Exception, match="needs encoder layer 3"):
        outs.hidden_states[-1]
    feats = get_spatio_temporal_features_torch(feats_in)
    assert feats.shape == (356, 1024) and feats.dtype == torch.float16
    ids = O.make_prompt_ids(lcfg, 356, seed=1).cuda()
    out = m.generate(ids, video_spatio_temporal_features=feats.unsqueeze(0), do_sample=False, max_new_tokens=6)
    assert out.shape == (1, 448 + 6) and out.dtype == torch.int64
    assert torch.equal(out[:, :448], ids)                           # prompt included, like HF generate
    bf = lambda sd: {k: v.cuda().bfloat16() for k, v in sd.items()}
    ref_toks, ref_logits = O.greedy_generate(bf(lsd), lcfg, ids, feats[None].bfloat16(), 6)
    lg = m(input_ids=ids, video_spatio_temporal_features=feats.unsqueeze(0)).logits
    assert lg.shape == (1, 1, lcfg.vocab)
    rel = ((lg[:, 0].float() - ref_logits[0]).norm() / ref_logits[0].norm()).item()
    assert rel < 3e-2, rel
    assert out[0, 448].item() == ref_toks[0, 0].item()
    # cached single-token step through forward(), as HF's generation loop drives it
    step = m(input_ids=out[:, 448:449], past_key_values=448, video_spatio_temporal_features=feats.unsqueeze(0)).logits
    assert step[:, 0].argmax(-1).item() == out[0, 449].item()

    class Tok:   # stop on token id == the third generated token
        def __init__(self, stop): self.stop = stop
        def __call__(self, text): return type("E", (), {"input_ids": [self.stop]})()
        def batch_decode(self, ids, skip_special_tokens=True): return [""]
    crit = KeywordsStoppingCriteria(["x"], Tok(int(out[0, 450])), ids)
    out2 = m.generate(ids, video_spatio_temporal_features=feats.unsqueeze(0), do_sample=False, max_new_tokens=6,
                      stopping_criteria=[crit])
    assert out2.shape[1] == 448 + 3 and torch.equal(out2, out[:, :451])
    out3 = m.generate(ids, video_spatio_temporal_features=feats.unsqueeze(0), do_sample=True, temperature=0.2,
                      max_new_tokens=4)
    assert out3.shape == (1, 452)
    bad = ids.clone(); bad[0, 64 + 357] = 5
    with pytest.raises(ValueError, match="video start tokens and video end tokens"):
        m.generate(bad, video_spatio_temporal_features=feats.unsqueeze(0), max_new_tokens=2)


@torch.no_grad()
def test_generate_continue_second_turn():
    """Second turn about the same video through the KV cache (generate_continue): the tokens it
    produces must be the ones a from-scratch generate() on the concatenated context produces."""
    lcfg = O.LlmCfg(hidden=512, inter=1024, heads=4, layers=2)
    m = _model(lcfg, 3, max_batch=1)
    m.load_state_dict(O.random_llm_state(lcfg, seed=23))
    feats = (torch.randn(356, 1024, generator=torch.Generator().manual_seed(3)) * 0.5).half().cuda()
    ids = O.make_prompt_ids(lcfg, 356, seed=2).cuda()
    turn1 = m.generate(ids, video_spatio_temporal_features=feats.unsqueeze(0), do_sample=False, max_new_tokens=5)
    q2 = torch.randint(3, 32000, (1, 11), generator=torch.Generator().manual_seed(8)).cuda()
    turn2 = m.generate_continue(q2, do_sample=False, max_new_tokens=5)
    assert turn2.shape == (1, 448 + 5 + 11 + 5)
    assert torch.equal(turn2[:, :453], turn1) and torch.equal(turn2[:, 453:464], q2)
    scratch = m.generate(torch.cat([turn1, q2], 1), video_spatio_temporal_features=feats.unsqueeze(0),
                         do_sample=False, max_new_tokens=5)
    # same arithmetic up to the GEMM tiling of the prefill: identical tokens except at bf16 near-ties
    assert (scratch[:, 464:] == turn2[:, 464:]).float().mean().item() >= 0.8
    assert scratch[0, 464].item() == turn2[0, 464].item()
    m2 = _model(lcfg, 3, max_batch=1)
    with pytest.raises(ValueError, match="no previous generate"):
        m2.generate_continue(q2)
