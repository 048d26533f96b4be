"""A tiny LOCAL checkpoint in the layout the reference's `initialize_model` loads
(video_chatgpt/eval/model_utils.py:82-150): a LLaVA-style model directory (config.json with
`mm_vision_tower` pointing at a CLIP directory, safetensors weights, tokenizer files) and a CLIP
directory (config.json, preprocessor_config.json, safetensors weights). Everything is synthetic:
seeded random weights from the oracle's generators, a WordLevel tokenizer that -- like LLaMA's --
prepends <s>, so "</s>" tokenizes to [bos, eos]."""
import json
import os

import torch

from oracle import vcl_oracle as O

VOCAB = 1000          # base vocabulary; the three video tokens become ids 1000..1002


def make_tiny_checkpoint(root, llm_layers=2, clip_layers=3, image=224):
    from safetensors.torch import save_file
    from tokenizers import Tokenizer, models, pre_tokenizers, processors
    from transformers import CLIPImageProcessor
    root = str(root)
    model_dir, clip_dir = os.path.join(root, "llava-tiny"), os.path.join(root, "clip-tiny")
    os.makedirs(model_dir), os.makedirs(clip_dir)

    # ---- CLIP directory ----
    ccfg = O.ClipCfg(hidden=1024, inter=1024, heads=16, layers=clip_layers, image=image)
    csd = O.random_clip_state(ccfg, seed=31)
    json.dump({"model_type": "clip_vision_model", "hidden_size": 1024, "intermediate_size": 1024,
               "num_hidden_layers": clip_layers, "num_attention_heads": 16, "image_size": image, "patch_size": 14,
               "layer_norm_eps": 1e-5, "hidden_act": "quick_gelu", "projection_dim": 768},
              open(os.path.join(clip_dir, "config.json"), "w"))
    save_file({k: v.to(torch.float16).contiguous() for k, v in csd.items()}, os.path.join(clip_dir, "model.safetensors"))
    CLIPImageProcessor(size={"shortest_edge": image}, crop_size={"height": image, "width": image}).save_pretrained(clip_dir)

    # ---- model directory ----
    lcfg = O.LlmCfg(hidden=512, inter=1024, heads=4, layers=llm_layers, vocab=VOCAB,
                    proj_type="linear" if image == 224 else "mlp2x_gelu")
    lsd = O.random_llm_state(lcfg, seed=41)
    cfg = {"model_type": "VideoChatGPT", "hidden_size": 512, "intermediate_size": 1024, "num_hidden_layers": llm_layers,
           "num_attention_heads": 4, "num_key_value_heads": 4, "vocab_size": VOCAB, "rms_norm_eps": 1e-5,
           "rope_theta": 10000.0, "max_position_embeddings": 2048, "bos_token_id": 1, "eos_token_id": 2, "pad_token_id": 0,
           "mm_vision_tower": clip_dir, "use_mm_proj": True, "mm_hidden_size": 1024}
    if image != 224:
        cfg["mm_projector_type"] = "mlp2x_gelu"
    json.dump(cfg, open(os.path.join(model_dir, "config.json"), "w"))
    save_file({k: v.to(torch.float16).contiguous() for k, v in lsd.items()}, os.path.join(model_dir, "model.safetensors"))
    vocab = {"<unk>": 0, "<s>": 1, "</s>": 2}
    for i in range(3, VOCAB):
        vocab[f"w{i}"] = i
    tok = Tokenizer(models.WordLevel(vocab, unk_token="<unk>"))
    tok.pre_tokenizer = pre_tokenizers.WhitespaceSplit()
    tok.post_processor = processors.TemplateProcessing(single="<s> $A", special_tokens=[("<s>", 1)])
    tok.save(os.path.join(model_dir, "tokenizer.json"))
    json.dump({"tokenizer_class": "PreTrainedTokenizerFast", "bos_token": "<s>", "eos_token": "</s>", "unk_token": "<unk>",
               "pad_token": "<unk>"}, open(os.path.join(model_dir, "tokenizer_config.json"), "w"))
    return dict(model_dir=model_dir, clip_dir=clip_dir, clip_cfg=ccfg, llm_cfg=lcfg,
                clip_sd={k: v.to(torch.float16) for k, v in csd.items()},
                llm_sd={k: v.to(torch.float16) for k, v in lsd.items()})
