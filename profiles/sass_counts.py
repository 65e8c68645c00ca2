#!/usr/bin/env python
"""SASS evidence per kernel: counts of the instructions that prove which hardware paths a kernel uses.
    python profiles/sass_counts.py > profiles/sass_r02.txt      (needs only cuobjdump; no GPU)
UTCHMMA / UTCQMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st (TMEM), UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk (TMA
bulk copy; .MULTICAST = cluster multicast), UTMALDG / UTMASTG = tensor-map TMA, SYNCS = mbarrier, UCGABAR = barrier.cluster,
FFMA2 = packed FP32x2 FMA, LDGSTS = cp.async, DFMA / DADD / DMUL = FP64."""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OPS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "UCGABAR", "FFMA2", "FFMA", "HFMA2",
       "LDGSTS", "LDS", "STS", "LDG", "STG", "MUFU", "DFMA", "DADD", "DMUL", "SHFL", "ATOM", "RED", "BAR"]
print("# cuobjdump -sass of pgmorl_b200/build/*.o (sm_100a), instruction counts per kernel (static)\n")
for obj in sorted(glob.glob(os.path.join(ROOT, "pgmorl_b200", "build", "*.o"))):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    arch = re.search(r"arch = (sm_\w+)", out)
    print(f"## {os.path.basename(obj)}  ({arch.group(1) if arch else '?'})")
    cur, counts, mc = None, collections.OrderedDict(), collections.Counter()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            cur = re.sub(r"\(pgm::\w+.*\)$|\(.*\)$", "", cur).replace("void pgm::", "")
            counts[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)(\.[A-Z0-9_.]+)?", line)
        if m and cur:
            op = m.group(1)
            for k in OPS:
                if op == k or (k == "UCGABAR" and op.startswith("UCGABAR")):
                    counts[cur][k] += 1
            if op == "UBLKCP" and m.group(2) and "MULTICAST" in m.group(2):
                counts[cur]["UBLKCP.MULTICAST"] += 1
    for k, c in counts.items():
        if sum(c.values()):
            print(f"{k}\n    " + "  ".join(f"{n}={v}" for n, v in c.items() if v))
    print()
