"""Debug driver: one decode step through gemv_tcw (run under compute-sanitizer)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "video-llava_b200"))
from oracle import vcl_oracle as O
from _util import make_engine, to_dev, vid_start_of
NB = int(sys.argv[1]) if len(sys.argv) > 1 else 5
import math
import vcl_native as vn
if os.environ.get("DBG_OP"):
    for (B, N, K) in [(5, 4096, 4096), (16, 4096, 11008), (9, 12288, 4096), (16, 32003, 4096)]:
        torch.manual_seed(B + N)
        x = torch.randn(B, K, device="cuda").bfloat16()
        w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).bfloat16()
        r = torch.randn(B, N, device="cuda").bfloat16()
        out = vn.op_gemv(x, w, r, None, 0.0)
        torch.cuda.synchronize()
        ref = (x.float() @ w.float().t()).bfloat16().float() + r.float()
        print("op_gemv", B, N, K, "rel", ((out.float() - ref).norm() / ref.norm()).item(), flush=True)
    sys.exit(0)
cfg = O.LlmCfg(hidden=2560, inter=6912, heads=20, layers=int(os.environ.get("DBG_LAYERS", "1")))
sd = O.random_llm_state(cfg, seed=8)
ids = O.make_prompt_ids(cfg, 356, seed=6, batch=NB).to("cuda")
vf = (torch.randn(NB, 356, 1024) * 0.5).half().float().to("cuda")
eng = make_engine(llm=cfg, max_batch=NB, max_seq=480)
eng.load_llm(to_dev(sd))
vs = vid_start_of(ids, cfg)
_, lg, _ = eng.prefill(ids, vf, vs, want_logits=True)
tok = lg.argmax(-1).to(torch.int32)
torch.cuda.synchronize(); print("prefill ok", flush=True)
lgb, tokb = eng.decode_step(tok, 448, want_logits=True)
torch.cuda.synchronize(); print("decode ok", tokb.tolist(), flush=True)
