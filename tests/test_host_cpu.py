"""CPU tests of the host side: the C-ABI library loads here and exports every symbol of
include/vcl.h (no compute without a GPU, and it must say so loudly), the reference-mirroring host
logic (frame sampling, prompt template, stop criterion, span validation) behaves like the reference,
and the world_size=2 sharding/gather path works over gloo."""
import ctypes
import os
import sys

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402


@pytest.fixture(scope="module")
def built():
    entry.build()
    import vcl_native
    return vcl_native


def test_cabi_exports_every_declared_symbol(built):
    lib = ctypes.CDLL(built.LIB_PATH)
    declared = entry.declared_symbols()
    assert len(declared) >= 15
    for s in declared:
        assert hasattr(lib, s), s
    assert sorted(built.EXPORTED_SYMBOLS) == declared
    assert built.lib().vcl_version() == 1


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback(built):
    h = ctypes.c_void_p()
    cfg = built.vcl_config()
    rc = built.lib().vcl_create(ctypes.byref(h), ctypes.byref(cfg))
    assert rc != 0 and not h.value
    assert b"no CPU fallback" in built.lib().vcl_last_error() or b"CUDA" in built.lib().vcl_last_error()
    with pytest.raises(built.VclError):
        built.check(built.lib().vcl_st_pool(ctypes.c_void_p(16), 1, 8, 8, 1, 1, 8, 100, ctypes.c_void_p(16), 1, None))
    from video_chatgpt.inference import get_spatio_temporal_features_torch
    with pytest.raises(built.VclError, match="GPU"):
        get_spatio_temporal_features_torch(torch.zeros(2, 4, 8, dtype=torch.float16))


def test_get_seq_frames_matches_reference_examples(built):
    from video_chatgpt.eval.model_utils import get_seq_frames
    assert get_seq_frames(1000, 100)[:4] == [5, 15, 25, 35]
    assert get_seq_frames(1000, 100)[-1] == 994
    assert get_seq_frames(37, 37) == list(range(37)) or len(get_seq_frames(37, 37)) == 37
    import numpy as np
    for total, want in [(250, 100), (101, 100), (8, 8), (3000, 100)]:
        seg = float(total - 1) / want
        ref = [(int(np.round(seg * i)) + int(np.round(seg * (i + 1)))) // 2 for i in range(want)]
        assert get_seq_frames(total, want) == ref


def test_prompt_template_and_stop_criterion(built):
    from video_chatgpt.model.utils import KeywordsStoppingCriteria
    from video_chatgpt.video_conversation import SeparatorStyle, conv_templates
    conv = conv_templates["pg-video-llava"].copy()
    conv.append_message(conv.roles[0], "What happens?")
    conv.append_message(conv.roles[1], None)
    p = conv.get_prompt()
    assert p.startswith("You are PG-Video-LLaVA, a large vision-language assistant. ")
    assert p.endswith(" USER: What happens? ASSISTANT:")
    assert conv.sep_style == SeparatorStyle.TWO and conv.sep2 == "</s>"
    assert conv_templates["pg-video-llava"].messages == []          # copy() does not alias

    class Tok:
        def __call__(self, text):
            return type("E", (), {"input_ids": [2] if text == "</s>" else [5, 6]})()

        def batch_decode(self, ids, skip_special_tokens=True):
            return ["".join(chr(97 + int(t) % 26) for t in row) for row in ids]

    prompt = torch.tensor([[1, 7, 8]])
    crit = KeywordsStoppingCriteria(["</s>"], Tok(), prompt)
    assert crit(torch.tensor([[1, 7, 8, 9]])) is False               # first call only records the length
    assert crit(torch.tensor([[1, 7, 8, 9, 4]])) is False
    assert crit(torch.tensor([[1, 7, 8, 9, 2]])) is True             # single-token keyword id
    crit2 = KeywordsStoppingCriteria(["jk"], Tok(), prompt)
    crit2(prompt)
    assert crit2(torch.tensor([[1, 7, 8, 9, 10]])) is True           # substring of the decoded text ("jk")


def test_video_feature_cache_lru(built):
    """The optional per-video cache of pooled features (multi-turn chats)."""
    from video_chatgpt.inference import VideoFeatureCache
    c = VideoFeatureCache(capacity=2)
    assert c.get("a") is None and c.misses == 1
    c.put("a", torch.zeros(1)); c.put("b", torch.ones(1))
    assert c.get("a") is not None and c.hits == 1              # "a" becomes the most recent
    c.put("c", torch.full((1,), 2.0))                           # evicts "b", the least recently used
    assert len(c) == 2 and c.get("b") is None and c.get("a") is not None and c.get("c") is not None
    with pytest.raises(ValueError):
        VideoFeatureCache(capacity=0)


def test_video_span_validation_errors(built):
    from oracle import vcl_oracle as O
    from video_chatgpt.model import VideoChatGPTConfig, VideoChatGPTLlamaForCausalLM
    cfg = VideoChatGPTConfig(hidden_size=512, intermediate_size=1024, num_hidden_layers=2, num_attention_heads=4,
                             vocab_size=32003, use_mm_proj=True, mm_hidden_size=1024)
    m = VideoChatGPTLlamaForCausalLM(cfg, clip_config={})
    vc = m.get_model().vision_config
    vc.vid_patch_token, vc.vid_start_token, vc.vid_end_token, vc.use_vid_start_end = 32000, 32001, 32002, True
    assert (vc.frame_size, vc.patch_size, vc.hidden_size) == (224, 14, 1024)
    lcfg = O.LlmCfg(hidden=512, inter=1024, heads=4, layers=2)
    ids = O.make_prompt_ids(lcfg, 356, seed=1, batch=2)
    assert m._video_spans(ids, 356) == [64, 64]
    bad = ids.clone(); bad[0, 64 + 357] = 5
    with pytest.raises(ValueError, match="number of video start tokens and video end tokens"):
        m._video_spans(bad, 356)
    shifted = ids.clone(); shifted[1, 64 + 357] = 32000; shifted[1, 64 + 358] = 32002
    with pytest.raises(ValueError, match="video end token should follow"):
        m._video_spans(shifted, 356)
    text_only = torch.randint(3, 1000, (1, 32))
    assert m._video_spans(text_only, 356) == [built.NO_VIDEO] and built.NO_VIDEO == -2 ** 31
    vc.use_vid_start_end = False
    assert m._video_spans(ids, 356) == [64, 64]                      # rows 65..420 replaced either way
    lead = torch.cat([torch.full((1, 356), 32000), torch.randint(3, 1000, (1, 8))], 1)
    assert m._video_spans(lead, 356) == [-1]                         # patch tokens from row 0: a real span, not "no video"
    with pytest.raises(ValueError, match="number of video patch tokens"):
        m._video_spans(ids, 300)
    # state handling: resize_token_embeddings pads with the mean row, strict load rejects unknown keys
    m.load_state_dict({"model.embed_tokens.weight": torch.arange(12.).view(4, 3), "lm_head.weight": torch.ones(4, 3)})
    m.resize_token_embeddings(6)
    assert m.state_dict()["model.embed_tokens.weight"].shape == (6, 3)
    assert torch.allclose(m.state_dict()["model.embed_tokens.weight"][5], torch.tensor([4.5, 5.5, 6.5]))
    with pytest.raises(RuntimeError, match="Unexpected"):
        m.load_state_dict({"vision_model.x": torch.zeros(1)}, strict=True)
    assert m.load_state_dict({"vision_model.x": torch.zeros(1)}, strict=False).unexpected_keys == ["vision_model.x"]
    kw = m.prepare_inputs_for_generation(ids, past_key_values=None, video_spatio_temporal_features="f")
    assert kw["input_ids"].shape == ids.shape and kw["video_spatio_temporal_features"] == "f"
    assert m.prepare_inputs_for_generation(ids, past_key_values=7)["input_ids"].shape == (2, 1)


def test_eos_defaults_and_finished_row_masking(built):
    """HF generate stops at config.eos_token_id implicitly: the mirror's defaults and the
    padding of finished rows, on CPU tensors."""
    from video_chatgpt.model import VideoChatGPTConfig, VideoChatGPTLlamaForCausalLM
    cfg = VideoChatGPTConfig(hidden_size=512, intermediate_size=1024, num_hidden_layers=2, num_attention_heads=4,
                             vocab_size=32003, use_mm_proj=True, mm_hidden_size=1024)
    m = VideoChatGPTLlamaForCausalLM(cfg, clip_config={})
    assert m._eos_pad("config", None) == (2, 2)                      # LLaMA default, pad falls back to eos
    assert m._eos_pad(None, None) == (None, None)                    # explicit None: fixed-length decoding
    cfg.eos_token_id, cfg.pad_token_id = 7, 0
    assert m._eos_pad("config", None) == (7, 0) and m._eos_pad(5, 9) == (5, 9)
    new = torch.tensor([[4, 7, 3, 3], [4, 5, 6, 8]])
    out, done = m._mask_finished(new, 7, 0)
    assert not done and out.tolist() == [[4, 7, 0, 0], [4, 5, 6, 8]]  # row 0 finished: padded after its EOS
    new = torch.tensor([[4, 7, 3, 3], [4, 5, 7, 8]])
    out, done = m._mask_finished(new, 7, 0)
    assert done and out.tolist() == [[4, 7, 0], [4, 5, 7]]            # all finished: cut at the longest row


def test_stop_string_cannot_catch_eos_with_a_bos_prepending_tokenizer(built, tmp_path):
    """With a LLaMA-like tokenizer "</s>" -> [bos, eos]: the stop criterion has no keyword id and the
    decoded text drops the special token, so only the EOS id passed to generate() can stop the loop."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from _checkpoint import make_tiny_checkpoint
    from transformers import AutoTokenizer
    from video_chatgpt.model.utils import KeywordsStoppingCriteria
    ck = make_tiny_checkpoint(tmp_path)
    tok = AutoTokenizer.from_pretrained(ck["model_dir"])
    assert tok("</s>").input_ids == [1, 2] and tok.eos_token_id == 2
    prompt = torch.tensor([[1, 7, 8]])
    crit = KeywordsStoppingCriteria(["</s>"], tok, prompt)
    assert crit.keyword_ids == []
    crit(prompt)
    assert crit(torch.tensor([[1, 7, 8, 9, 2]])) is False            # EOS generated, criterion blind to it


def test_training_side_reader_round_trip(built, tmp_path):
    """The on-disk format of the extractor (one pickle of a [100+P,1024] float16 ndarray per video) as
    the reference's trainer consumes it (train/train.py:401-405, 447-452)."""
    import pickle
    import numpy as np
    from video_chatgpt.train import collate_video_features, load_video_features
    a = np.random.default_rng(0).standard_normal((356, 1024)).astype(np.float16)
    b = np.random.default_rng(1).standard_normal((356, 1024)).astype(np.float16)
    for name, arr in (("v1.pkl", a), ("v2.pkl", b)):
        with open(tmp_path / name, "wb") as f:
            pickle.dump(arr, f)
    fa, fb = load_video_features(str(tmp_path), "v1.pkl"), load_video_features(str(tmp_path), "v2.pkl")
    assert isinstance(fa, np.ndarray) and fa.dtype == np.float16 and fa.shape == (356, 1024)
    batch = collate_video_features([{"video": fa}, {"video": fb}])
    assert batch.shape == (2, 356, 1024) and batch.dtype == torch.float16
    assert np.array_equal(batch[1].numpy(), b)
    ragged = collate_video_features([{"video": fa}, {"video": fb[:300]}])
    assert isinstance(ragged, list) and ragged[1].shape == (300, 1024)


def _dp_worker(rank, world, port, n_clips, ret):
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "video-llava_b200"))
    import dp
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    mine = dp.shard_clips(n_clips, rank, world)
    local = torch.tensor([[c * 100 + j for j in range(4)] for c in mine], dtype=torch.int32).reshape(len(mine), 4)
    full = dp.gather_tokens(local, n_clips, rank, world, dist)
    ret[rank] = full.tolist()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_clips", [5, 4, 1])
def test_sharding_and_gather_world2_gloo(n_clips):
    sys.path.insert(0, os.path.join(ROOT, "video-llava_b200"))
    import dp
    owned = sorted(dp.shard_clips(n_clips, 0, 2) + dp.shard_clips(n_clips, 1, 2))
    assert owned == list(range(n_clips))
    with pytest.raises(ValueError):
        dp.shard_clips(4, 2, 2)
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000) + n_clips
    mp.spawn(_dp_worker, args=(2, port, n_clips, ret), nprocs=2, join=True)
    want = [[c * 100 + j for j in range(4)] for c in range(n_clips)]
    assert ret[0] == want and ret[1] == want
