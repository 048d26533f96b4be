"""DRAM traffic of one decode step from an `ncu --set full` capture of tools/profile_step.py (eager decode
steps, PROF_LLM_LAYERS layers): sums dram__bytes_read + dram__bytes_write over the kernels of the LAST captured
step and scales the per-layer part to the full depth. Writes profiles/r02_decode_step_traffic.json, which
bench.py reports as roofline.traffic (x the 31 steps of the loop).

    python tools/decode_traffic.py gpurun_out/ncu_full_<tag>.csv <layers captured> [model layers = 32]
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path, L = sys.argv[1], int(sys.argv[2])
L_full = int(sys.argv[3]) if len(sys.argv) > 3 else 32
rows = list(csv.reader(open(path)))
hdr, units = rows[0], rows[1]
ir, iw, it = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
k = [(r[0].split("(")[0].replace("void ", "").replace("vcl::", "").replace("<unnamed>::", ""),
      float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]], float(r[it])) for r in rows[2:] if len(r) > it and r[ir]]
per_step = 5 * L + 2                      # (q|k|v, attention, o_proj, gate/up, down) x L + logits + arg-max
# the capture window may start anywhere in the step sequence: anchor on the last arg-max launch
ia = max(i for i, r in enumerate(k) if "argmax" in r[0])
head = k[ia - 1:ia + 1]
body = k[ia + 1:ia + 1 + 5 * L] if len(k) - ia - 1 >= 5 * L else k[ia - 1 - 5 * L:ia - 1]
assert len(body) == 5 * L and "gemv" in head[0][0] and sum("attn" in r[0] for r in body) == L, [r[0] for r in body]
step = body + head
layer_bytes = sum(b for _, b, _ in body) / L
head_bytes = head[0][1] + head[1][1]
out = {"source": os.path.relpath(path, ROOT), "layers_captured": L, "model_layers": L_full,
       "per_layer_dram_bytes": layer_bytes, "head_dram_bytes": head_bytes,
       "step_dram_bytes": layer_bytes * L_full + head_bytes,
       "kernels_of_the_step": [{"kernel": n, "dram_bytes": b, "us": t} for n, b, t in step]}
json.dump(out, open(os.path.join(ROOT, "profiles", "r02_decode_step_traffic.json"), "w"), indent=1)
print(json.dumps({k2: v for k2, v in out.items() if k2 != "kernels_of_the_step"}))
