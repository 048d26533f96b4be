"""Per-kernel SASS opcode histogram of the built libvcl.so (cuobjdump -sass): which kernels are
genuinely tcgen05 / TMA / TMEM (UTCHMMA, UTMALDG, UTMASTG, LDTM, UBLKCP) and which use mma.sync (HMMA)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "video-llava_b200", "libvcl.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
OPS = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UTMAPF", "LDTM", "STTM", "UBLKCP", "UTCBAR", "HMMA", "MUFU", "LDGSTS", "SYNCS"]
cur, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(anonymous namespace\)::", "", cur)
        cur = re.sub(r"\((?:[^()]|\([^()]*\))*\)$", "", cur).replace("void ", "")
        hist[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m:
        hist[cur]["_total"] += 1
        op = m.group(1)
        for o in OPS:
            if op.startswith(o):
                hist[cur][o] += 1
print(f"# cuobjdump -sass {os.path.relpath(lib, ROOT)}  (sm_100a); counts of instructions per kernel")
print("kernel | total | " + " | ".join(OPS))
tot = collections.Counter()
for k, c in hist.items():
    print(f"{k} | {c['_total']} | " + " | ".join(str(c[o]) for o in OPS))
    tot.update(c)
print("ALL | " + str(tot["_total"]) + " | " + " | ".join(str(tot[o]) for o in OPS))
