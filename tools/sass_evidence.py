#!/usr/bin/env python
"""profiles/sass_evidence_rNN.txt: per-kernel counts of the SASS mnemonics that prove the Blackwell paths
(cuobjdump -sass of the built library; needs no GPU)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "p2pvg_b200", "libp2pvg_b200.so")
out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "sass_evidence_r02.txt")
KEYS = ["UTCHMMA", "UTMALDG", "LDTM", "STTM", "UTCBAR", "HMMA", "UCGABAR", "MAPA"]
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
counts, cur = collections.OrderedDict(), None
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m:
        op = m.group(1)
        for k in KEYS:
            if op.startswith(k):
                counts[cur][k] += 1
names = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
rows = []
for mangled, name in zip(counts, names):
    c = counts[mangled]
    if not c:
        continue
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*", "", name)
    rows.append((name, "  ".join(f"{k}={c[k]}" for k in KEYS if c[k])))
rows.sort()
with open(out, "w") as f:
    f.write("# cuobjdump -sass p2pvg_b200/libp2pvg_b200.so (sm_100a): instruction counts per kernel (tools/sass_evidence.py)\n"
            "#   UTCHMMA = tcgen05.mma, UTMALDG = TMA tensor load (cp.async.bulk.tensor), LDTM / STTM = tcgen05.ld / tcgen05.st (tensor memory <-> registers),\n"
            "#   UTCBAR = tcgen05.commit -> mbarrier, HMMA = mma.sync m16n8k8 tf32 (LSTM scans), UCGABAR = barrier.cluster arrive / wait,\n"
            "#   MAPA = mapa.shared::cluster (distributed shared memory)\n\n")
    for name, c in rows:
        f.write(f"{name[:78]:78s} {c}\n")
print(f"{len(rows)} kernels -> {out}")
