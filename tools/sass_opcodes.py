#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of the in-tree library (cuobjdump -sass), for profiles/: which kernels carry bulk-TMA
(UBLKCP), tensor-map TMA (UTMASTG / UTMALDG), mbarrier (SYNCS), cp.async (LDGSTS), warp reductions (REDUX), and how
many FP64 / LSU instructions.   python tools/sass_opcodes.py > profiles/r05_sass_opcodes.txt"""
import collections, re, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
so = ROOT / "qrusty_b200" / "lib" / "libqrusty_cuda.so"
txt = subprocess.run(["cuobjdump", "-sass", str(so)], capture_output=True, text=True, check=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", txt)))
kernels, cur = collections.OrderedDict(), None
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        kernels[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and cur:
        kernels[cur][m.group(1)] += 1
KEY = ["UBLKCP", "UTMASTG", "UTMALDG", "SYNCS", "LDGSTS", "REDUX", "DFMA", "DADD", "POPC", "LDG", "STG", "LDS", "STS", "BAR", "ATOMS", "UCGABAR_ARV"]
print("library:", so.name, " cubin arch:", ", ".join(arch), " kernels:", len(kernels))
print("UBLKCP = cp.async.bulk (1-D TMA), UTMASTG / UTMALDG = cp.async.bulk.tensor store / load (tensor-map TMA), SYNCS = mbarrier,")
print("LDGSTS = cp.async, REDUX = warp reduce, UCGABAR_ARV = cluster barrier.  No UTC*MMA / HMMA anywhere: the path has no contraction.\n")
print("%-62s %6s  %s" % ("kernel", "instr", "  ".join("%s" % k for k in KEY)))
tot = collections.Counter()
for name, c in kernels.items():
    tot.update(c)
    print("%-62s %6d  %s" % (name[:62], sum(c.values()), "  ".join("%*d" % (len(k), c.get(k, 0)) for k in KEY)))
print("\nall kernels: " + ", ".join("%s %d" % (k, tot[k]) for k in KEY))
mma = [k for k in tot if "MMA" in k]
print("tensor-core opcodes:", mma if mma else "none")
