"""Timeline of the per-matrix decode kernels (gemv_tc.cu) over one eager decode step (VCL_TC_TRACE)."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "video-llava_b200"))
os.environ["VCL_TC_TRACE"] = "1"
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
path = os.path.join(ROOT, "gpurun_out", "tc_trace.bin")
lib = vn.lib()
lib.vcl_debug_tc_trace_dump.argtypes = [ctypes.c_char_p]
print("dump rc", lib.vcl_debug_tc_trace_dump(path.encode()))
raw = np.fromfile(path, dtype=np.uint64)
n_rec, G = int(raw[0]), int(raw[1])
t = raw[2:].reshape(-1, G, 4, 8).astype(np.int64)
fused = os.environ.get("VCL_DECODE_FUSED") is not None
per_step = (L + 1) if fused else (4 * L + 1)
recs = t[n_rec - per_step:n_rec]
names = {2: "qkv", 0: "res", 1: "swiglu", 3: "logits"}
t0 = recs[0][:, 0, 0].min()
print("fused" if fused else "unfused", "step span us", (recs[-1][:, :, 4].max() - t0) / 1e3, "launches", per_step)
if fused:
    # launches 1..L: [o_proj, gate/up, down, next q|k|v or logits]
    for ph in range(4):
        rows = []
        for k in range(2, L):            # layers 1..L-2 (phase 3 = q|k|v)
            r = recs[k][:, ph, :]
            prev_end = recs[k][:, ph - 1, 4] if ph > 0 else None
            d = dict(x_stage=(r[:, 2] - r[:, 1]).mean() / 1e3, stream=(r[:, 3] - r[:, 2]).mean() / 1e3,
                     stream_max=(r[:, 3].max() - r[:, 2].min()) / 1e3, epi=(r[:, 4] - r[:, 3]).mean() / 1e3,
                     barrier=((r[:, 1] - prev_end).mean() / 1e3 if ph > 0 else 0.0),
                     barrier_last_arrival_to_release=((r[:, 1].min() - prev_end.max()) / 1e3 if ph > 0 else 0.0),
                     phase_total=(r[:, 4].max() - (prev_end.max() if ph > 0 else r[:, 1].min())) / 1e3,
                     producer_ahead=(r[:, 3] - r[:, 6]).mean() / 1e3)
            rows.append(d)
        print("phase", ph, names[int(recs[2][0, ph, 7] >> 32)], " ".join("%s=%.2f" % (kk, float(np.mean([d[kk] for d in rows]))) for kk in rows[0]))
    lay = [(recs[k + 1][:, 0, 1].min() - recs[k][:, 0, 1].min()) / 1e3 for k in range(1, L - 1)]
    gaps = [(recs[k + 1][:, 0, 1].min() - recs[k][:, 3, 4].max()) / 1e3 for k in range(1, L - 1)]
    print("per-layer us", [round(x, 1) for x in lay])
    print("fused-kernel end -> next fused kernel's dependency resolved (attention in between) us", [round(x, 1) for x in gaps])
else:
    for k in range(4, per_step, max(1, (per_step - 4) // 8)):
        r = recs[k][:, 0, :]
        print(k, names[int(r[0, 7] >> 32)], "x_stage %.2f stream %.2f/%.2f epi %.2f" % ((r[:, 2] - r[:, 1]).mean() / 1e3, (r[:, 3] - r[:, 2]).mean() / 1e3,
              (r[:, 3].max() - r[:, 2].min()) / 1e3, (r[:, 4] - r[:, 3]).mean() / 1e3))
    lay = [(recs[4 * (l + 1)][:, 0, 1].min() - recs[4 * l][:, 0, 1].min()) / 1e3 for l in range(L - 1)]
    print("per-layer us", [round(x, 1) for x in lay])
sys.stdout.flush()
os._exit(0)
