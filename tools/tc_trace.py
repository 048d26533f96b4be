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
t = raw[2:].reshape(-1, G, 8).astype(np.int64)
per_step = 4 * L + 1
recs = t[n_rec - per_step:n_rec]
names = {2: "qkv", 0: "res", 1: "swiglu", 3: "logits"}
mb = {("qkv", 12288): 100.7, ("res", 4096): None, ("swiglu", 22016): 180.4, ("logits", 32003): 262.2}
t0 = recs[0][:, 0].min()
print("step span us", (recs[-1][:, 4].max() - t0) / 1e3, "kernels", per_step)
rows = {}
prev_end = None
for k in range(per_step):
    r = recs[k]
    mode = int(r[0, 7] >> 32); N = int(r[0, 7] & 0xffffffff)
    key = names[mode] + ("_o" if (mode == 0 and k % 4 == 1) else "_down" if mode == 0 else "")
    entry, waitd, xst, loop, end, p0, p1 = (r[:, i] for i in range(7))
    d = dict(entry_min=(entry.min() - t0) / 1e3, entry_spread=(entry.max() - entry.min()) / 1e3,
             wait_done=(waitd.min() - t0) / 1e3, wait_spread=(waitd.max() - waitd.min()) / 1e3,
             x_stage=(xst - waitd).mean() / 1e3, stream=(loop - xst).mean() / 1e3, stream_max=(loop.max() - xst.min()) / 1e3,
             epi=(end - loop).mean() / 1e3, total=(end.max() - waitd.min()) / 1e3,
             prefetch_lead=(waitd - p0).mean() / 1e3, prod_done_before_end=(end - p1).mean() / 1e3,
             gap_prev_end_to_wait=((waitd.min() - prev_end) / 1e3 if prev_end is not None else 0.0),
             end=(end.max() - t0) / 1e3)
    prev_end = end.max()
    if k >= 4:
        rows.setdefault(key, []).append(d)
for key, ds in rows.items():
    avg = {kk: float(np.mean([d[kk] for d in ds])) for kk in ds[0] if kk not in ("entry_min", "wait_done", "end")}
    print(key, " ".join("%s=%.2f" % kv for kv in avg.items()))
apath = os.path.join(ROOT, "gpurun_out", "attn_trace.bin")
has_attn = hasattr(lib, "vcl_debug_attn_trace_dump")   # only in builds with the attention trace hook
if has_attn and lib.vcl_debug_attn_trace_dump(apath.encode()) == 0:
    ar = np.fromfile(apath, dtype=np.uint64)
    an, ac = int(ar[0]), int(ar[1])
    at = ar[2:].reshape(-1, ac, 8).astype(np.int64)[an - L:an]          # last step: one launch per layer
    rows_a = []
    for l in range(1, L):
        a = at[l]
        qkv_end = recs[4 * l][:, 4].max(); o_wait = recs[4 * l + 1][:, 1].min()
        rows_a.append(dict(entry_before_qkv_end=(qkv_end - a[:, 0].max()) / 1e3, qkv_end_to_wait=(a[:, 1].min() - qkv_end) / 1e3,
                           wait_to_q=(a[:, 4] - a[:, 1]).mean() / 1e3, q_to_scores=(a[:, 5] - a[:, 4]).mean() / 1e3,
                           scores_to_sync1=(a[:, 2] - a[:, 5]).mean() / 1e3, sync1=(a[:, 6] - a[:, 2]).mean() / 1e3,
                           pv=(a[:, 7] - a[:, 6]).mean() / 1e3, sync2_write=(a[:, 3] - a[:, 7]).mean() / 1e3,
                           total_after_wait=(a[:, 3].max() - a[:, 1].min()) / 1e3, end_to_o_wait=(o_wait - a[:, 3].max()) / 1e3))
    print("attention", " ".join("%s=%.2f" % (k, float(np.mean([r[k] for r in rows_a]))) for k in rows_a[0]))
lay = [(recs[4 * (l + 1)][:, 1].min() - recs[4 * l][:, 1].min()) / 1e3 for l in range(L - 1)]
print("per-layer us", [round(x, 1) for x in lay])
sys.stdout.flush()
os._exit(0)
