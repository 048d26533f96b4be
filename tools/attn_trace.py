"""Per-CTA phase timeline of the ViT attention kernel (globaltimer stamps; debug hook vcl_debug_set_attn_trace)."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "video-llava_b200"))
import vcl_native as vn  # noqa: E402

N, S, H = 100, 257, 16
dev = torch.device("cuda:0")
qkv = (torch.randn(N * S, 3 * H * 64, device=dev) * 0.7).to(torch.bfloat16)
for _ in range(3):
    vn.op_attention_vit(qkv, N, S, H)
persistent = os.environ.get("VCL_ATTN_ONE_SHOT") is None
lib = vn.lib()
lib.vcl_debug_set_attn_trace.argtypes = [ctypes.c_void_p]
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
if persistent:
    G = 2 * torch.cuda.get_device_properties(0).multi_processor_count
    buf = torch.zeros(G * 12 * 8, dtype=torch.int64, device=dev)
    lib.vcl_debug_set_attn_trace(ctypes.c_void_p(buf.data_ptr()))
    s.record(); vn.op_attention_vit(qkv, N, S, H); e.record()
    torch.cuda.synchronize()
    lib.vcl_debug_set_attn_trace(ctypes.c_void_p(0))
    t = buf.cpu().numpy().reshape(G, 12, 8).astype(np.int64)[:, 1:10]      # tiles 1..9 of every CTA (steady state)
    names = ["wait for Q|K (fetched during the previous epilogue) + row-256 vectors", "dot64 for key / query 256, wait for S",
             "pass 1: TMEM -> bf16 stash, row max, block barrier", "pass 2: exp2, P -> smem", "wait for P.V", "epilogue (TMEM -> global)"]
    print(f"persistent kernel {s.elapsed_time(e) * 1e3:.1f} us; per tile mean {(t[:, :, 6] - t[:, :, 0]).mean() / 1e3:.2f} us; "
          f"tile-to-tile {(t[:, 1:, 0] - t[:, :-1, 0]).mean() / 1e3:.2f} us")
    for i, nm in enumerate(names):
        d = (t[:, :, i + 1] - t[:, :, i]) / 1e3
        print(f"  {nm:70s} mean {d.mean():6.2f} us  p10 {np.percentile(d, 10):6.2f}  p90 {np.percentile(d, 90):6.2f}")
    sys.exit(0)
n_cta = N * H * 2
buf = torch.zeros(n_cta * 8, dtype=torch.int64, device=dev)
lib.vcl_debug_set_attn_trace(ctypes.c_void_p(buf.data_ptr()))
s.record(); vn.op_attention_vit(qkv, N, S, H); e.record()
torch.cuda.synchronize()
lib.vcl_debug_set_attn_trace(ctypes.c_void_p(0))
t = buf.cpu().numpy().reshape(n_cta, 8).astype(np.int64)
names = ["setup (launch -> tmem alloc + barrier init)", "TMA load of Q, K, V (80 KB)", "S = Q.K^T (4 MMAs + commit)",
         "pass 1: TMEM -> bf16 stash, row max, block barrier", "pass 2: exp2, P -> smem", "P.V (16 MMAs + commit)", "epilogue + teardown"]
print(f"kernel {s.elapsed_time(e) * 1e3:.1f} us for {n_cta} CTAs; per-CTA lifetime mean {(t[:, 7] - t[:, 0]).mean() / 1e3:.2f} us; "
      f"span {(t[:, 7].max() - t[:, 0].min()) / 1e3:.1f} us")
for i, nm in enumerate(names):
    d = (t[:, i + 1] - t[:, i]) / 1e3
    print(f"  {nm:55s} mean {d.mean():6.2f} us  p10 {np.percentile(d, 10):6.2f}  p90 {np.percentile(d, 90):6.2f}")
