"""Timeline of the first k-blocks of the cta_group::2 GEMM (debug hook vcl_debug_set_gemm_trace)."""
import ctypes, math, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "video-llava_b200"))
import vcl_native as vn
dev = torch.device("cuda:0")
M, N, K = [int(x) for x in os.environ.get("PAIR_SHAPE", "256,256,8192").split(",")]   # e.g. PAIR_SHAPE=25700,3072,1024
a = torch.randn(M, K, device=dev).bfloat16(); w = (torch.randn(N, K, device=dev) / math.sqrt(K)).bfloat16()
for _ in range(3): vn.op_gemm(a, w, None, None, vn.ACT_NONE, 256, cluster=-2)
buf = torch.zeros(2 * 128 * 4, dtype=torch.int64, device=dev)
lib = vn.lib(); lib.vcl_debug_set_gemm_trace.argtypes = [ctypes.c_void_p]
lib.vcl_debug_set_gemm_trace(ctypes.c_void_p(buf.data_ptr()))
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record(); vn.op_gemm(a, w, None, None, vn.ACT_NONE, 256, cluster=-2); e.record(); torch.cuda.synchronize()
lib.vcl_debug_set_gemm_trace(ctypes.c_void_p(0))
s2, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s2.record(); vn.op_gemm(a, w, None, None, vn.ACT_NONE, 256, cluster=1); e2.record(); torch.cuda.synchronize()
t = buf.cpu().numpy().reshape(2, 128, 4).astype(np.int64)
t0 = t[0, 0, 0]
print(f"pair kernel {s.elapsed_time(e)*1e3:.1f} us for {K//64} k-blocks; single-CTA kernel (2 CTAs) {s2.elapsed_time(e2)*1e3:.1f} us")
print("kb | leader producer: empty-wait start, ready | peer producer: start, ready | leader MMA: full-wait start, ready   (ns from start)")
for kb in list(range(0, 24)) + [60, 100, 127]:
    print(kb, t[0, kb, 0] - t0, t[0, kb, 1] - t0, "|", t[1, kb, 0] - t0, t[1, kb, 1] - t0, "|", t[0, kb, 2] - t0, t[0, kb, 3] - t0)
nk = K // 64
if nk < 128:
    print(f"tile boundaries every {nk} k-blocks: ns between consecutive full-barrier completions around them")
    dd = np.diff(t[0, :128, 3])
    for b in range(nk, 120, nk):
        print("  k-block", b, ":", dd[b - 3:b + 3].tolist(), "| MMA waited", (t[0, b, 3] - t[0, b, 2]), "ns for its operands, issue gap to the previous", t[0, b, 2] - t[0, b - 1, 3])
d = np.diff(t[0, 8:120, 3])
print("steady state: ns between consecutive full-barrier completions: mean", d.mean(), "p10", np.percentile(d, 10), "p90", np.percentile(d, 90))
