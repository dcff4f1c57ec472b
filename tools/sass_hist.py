#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of the built library (cuobjdump -sass): the tracked evidence for which instructions the hot kernels
are made of (DMMA.8x8x4 = FP64 tensor core, UBLKCP = cp.async.bulk / TMA bulk copy, SYNCS = mbarrier, LDS/STS, ATOMS ...).
usage: python tools/sass_hist.py [lib.so] > profiles/rNN_sass_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "abcsmc_b200", "libabcsmc_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
kern, hist, order = None, {}, []
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = m.group(1); hist[kern] = collections.Counter(); order.append(kern); continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Za-z0-9_.]*)", line)
    if m and kern:
        op = m.group(1)
        base = op.split(".")[0]
        key = op if base in ("DMMA", "UBLKCP", "SYNCS", "ATOMS", "ATOMG", "RED", "LDGSTS", "UTMALDG", "BAR", "SHFL", "F2I", "F2F", "MUFU") else base
        hist[kern][key] += 1
print(f"# {os.path.basename(lib)}: {len(order)} kernels; arch {re.findall(r'arch = (sm_[0-9a-z]+)', txt)[:1]}")
tot = collections.Counter()
for k in order:
    tot.update(hist[k])
print("# whole library:", ", ".join(f"{o}={c}" for o, c in tot.most_common(40)))
keys = ["DMMA.8x8x4", "DFMA", "DADD", "DMUL", "UBLKCP.S.G", "SYNCS", "LDS", "STS", "ATOMS", "LDG", "STG", "SHFL", "BAR", "F2I", "MUFU"]
for k in sorted(order, key=lambda k: -sum(hist[k].values())):
    h = hist[k]
    name = demangle(k)
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    n = sum(h.values())
    sel = []
    for key in keys:
        c = sum(v for o, v in h.items() if o == key or o.startswith(key + ".") or (key == "SYNCS" and o.startswith("SYNCS")))
        if c:
            sel.append(f"{key}={c}")
    print(f"{name[:110]:110s} {n:6d} instr | " + " ".join(sel))
