#!/usr/bin/env python
"""SASS digest of libb2b.so (run here, no GPU): per kernel, the count of the mnemonics that prove a Blackwell-native kernel
(B200_PROFILING.md): UTMALDG / UTMASTG / UBLKCP (TMA), UTC*MMA (tcgen05.mma), LDTM / STTM (tcgen05.ld/st), FFMA2 (packed
fp32), plus MUFU / LDS / STS / LDG / STG / LDL / STL and the instruction total.  python tools/sass_digest.py > profiles/r02_sass_digest.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "bijectors.jl_b200", "csrc", "libb2b.so")
KEYS = ["UTMALDG", "UTMASTG", "UBLKCP", "UTC", "LDTM", "STTM", "FFMA2", "FFMA", "MUFU", "LDS", "STS", "LDG", "STG", "LDL", "STL", "REDUX", "SHFL"]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n  # noqa: E731
cur, counts, total = None, collections.defaultdict(collections.Counter), collections.Counter()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        total[cur] += 1
        for k in KEYS:
            if op.split(".")[0] == k or (k == "UTC" and op.startswith("UTC") and "MMA" in op):
                counts[cur][k] += 1
print(f"# SASS digest of {os.path.relpath(lib, ROOT)} (cuobjdump -sass); columns: " + " ".join(KEYS) + " | total")
agg = collections.Counter()
for fn in sorted(total, key=lambda f: demangle(f)):
    name = demangle(fn)
    cut = name.rfind(">(")
    name = name[: cut + 1] if cut > 0 else re.sub(r"\(.*", "", name)
    name = name.replace("void ", "").replace("b2b::", "").replace("(int)", "").replace("(bool)", "")
    print(f"{name[:78]:78s} " + " ".join(f"{counts[fn][k]:5d}" for k in KEYS) + f" | {total[fn]:6d}")
    for k in KEYS:
        agg[k] += counts[fn][k]
print(f"{'ALL KERNELS':78s} " + " ".join(f"{agg[k]:5d}" for k in KEYS) + f" | {sum(total.values()):6d}")
