"""Timeline of one decode step of the persistent megakernel (VCL_MEGAKERNEL=1, VCL_MEGA_TRACE=<file>)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "video-llava_b200"))
path = os.path.join(ROOT, "gpurun_out", "mega_trace.bin")
os.environ["VCL_MEGAKERNEL"] = "1"
os.environ["VCL_MEGA_TRACE"] = path
import bench  # noqa: E402
import vcl_native as vn  # noqa: E402

L = int(os.environ.get("PROF_LLM_LAYERS", "8"))
bench.MODELS["7b"]["layers"] = L
dev = torch.device("cuda:0")
c = vn.vcl_config()
c.clip_layers, c.clip_hidden, c.clip_inter, c.clip_heads = 1, 1024, 4096, 16
c.image_size, c.patch_size, c.clip_ln_eps = 224, 14, 1e-5
c.llm_layers, c.llm_hidden, c.llm_inter, c.llm_heads = L, 4096, 11008, 32
c.vocab, c.rms_eps, c.rope_theta = 32003, 1e-5, 10000.0
c.proj_type, c.n_temporal = vn.PROJ_LINEAR, 100
c.max_frames, c.max_batch, c.max_seq = 100, 1, 480
eng = vn.Engine(c)
_, llm_sd = bench.device_weights("7b", dev)
eng.load_llm(llm_sd)
del llm_sd
ids = torch.randint(3, 32000, (1, 448), device=dev); ids[:, 64] = 32001; ids[:, 65:421] = 32000; ids[:, 421] = 32002
vs = torch.full((1,), 64, dtype=torch.int32, device=dev)
feats = torch.randn(1, 356, 1024, device=dev).to(torch.bfloat16)
_, _, tok = eng.prefill(ids, feats, vs)
for i in range(6):
    _, tok = eng.decode_step(tok, 448 + i)
torch.cuda.synchronize()
t = np.fromfile(path, dtype=np.uint64).reshape(-1, L * 6 + 1, 8).astype(np.int64)
G = t.shape[0]
t0 = t[:, 0, 0].min()
print("CTAs", G, "step total us", (t[:, L * 6, 4].max() - t0) / 1e3)
names = ["qkv", "attA", "attB", "o", "gu", "down"]
ideal = {"qkv": 100.7, "o": 33.6, "gu": 180.4, "down": 90.2}
for ph in range(6):
    rows = []
    for l in range(1, L):
        r = t[:, l * 6 + ph, :]
        start = r[:, 0]; sync = r[:, 4]
        glob_start = start.min()
        d = {"start_spread": (start.max() - start.min()) / 1e3,
             "x": np.where(r[:, 1] > 0, r[:, 1] - start, 0).mean() / 1e3,
             "mma_mean": np.where(r[:, 2] > 0, r[:, 2] - np.maximum(r[:, 1], start), 0).mean() / 1e3,
             "mma_max": np.where(r[:, 2] > 0, r[:, 2] - np.maximum(r[:, 1], start), 0).max() / 1e3,
             "work_end_max": (r[:, 3].max() - glob_start) / 1e3,
             "work_end_min": (r[:, 3][r[:, 3] > 0].min() - glob_start) / 1e3,
             "sync_end": (sync.max() - glob_start) / 1e3,
             "prod_done_mean": (np.where(r[:, 5] > 0, r[:, 5] - glob_start, 0).mean()) / 1e3}
        rows.append(d)
    keys = rows[0].keys()
    avg = {k: float(np.mean([r[k] for r in rows])) for k in keys}
    extra = ""
    if names[ph] in ideal:
        extra = " ideal@6574GB/s %.1f us" % (ideal[names[ph]] / 6574 * 1e3)
    print(names[ph], " ".join("%s=%.2f" % (k, v) for k, v in avg.items()) + extra)
r = t[:, L * 6, :]
print("lm_head total us", (r[:, 4].max() - r[:, 0].min()) / 1e3, "ideal", 262.2 / 6574 * 1e3)
lay = [(t[:, (l + 1) * 6, 0].min() - t[:, l * 6, 0].min()) / 1e3 for l in range(L - 1)]
print("per-layer us", [round(x, 1) for x in lay])
sys.stdout.flush()
os._exit(0)
