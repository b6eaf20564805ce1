#!/usr/bin/env python
"""tools/sass_summary.py -- SASS evidence for profiles/: per kernel of libbev_b200.so, how many TMA /
mbarrier / dp2a ... instructions the sm_100a code holds (mnemonics: B200_PROFILING.md)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "bev_b200", "libbev_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
cols = ["UTMALDG", "UTMAPF", "UBLKCP", "SYNCS", "IDP", "LDS", "STG", "LDG", "SHFL", "FP64"]
pat = {"UTMALDG": r"\bUTMALDG", "UTMAPF": r"\bUTMAPF", "UBLKCP": r"\bUBLKCP", "SYNCS": r"\bSYNCS",
       "IDP": r"\bIDP\.", "LDS": r"\bLDS\b", "STG": r"\bSTG\b", "LDG": r"\bLDG\b", "SHFL": r"\bSHFL\b",
       "FP64": r"\b(DFMA|DMUL|DADD)\b"}
kern, rows, arch = None, {}, set()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = m.group(1)
        rows[kern] = collections.Counter()
        continue
    m = re.search(r"arch = (sm_\w+)", line)
    if m:
        arch.add(m.group(1))
    if kern and re.search(r"/\*[0-9a-f]{4,5}\*/", line):
        rows[kern]["n"] += 1
        for c in cols:
            if re.search(pat[c], line):
                rows[kern][c] += 1
names = subprocess.run(["c++filt"], input="\n".join(rows), capture_output=True, text=True).stdout.splitlines()
print("# cuobjdump -sass bev_b200/libbev_b200.so: instruction counts per kernel, arch %s" % ",".join(sorted(arch)))
print("# UTMALDG = cp.async.bulk.tensor (TMA tensor load), UTMAPF = TMA L2 prefetch, UBLKCP = cp.async.bulk,")
print("# SYNCS = mbarrier operations, IDP = dp2a / dp4a")
print("kernel | instructions | " + " | ".join(cols))
for k, nm in sorted(zip(rows, names), key=lambda t: t[1]):
    nm = re.sub(r"\(anonymous namespace\)::", "", nm)
    nm = re.sub(r"\(.*", "", nm)
    print("%s | %d | %s" % (nm, rows[k]["n"], " | ".join(str(rows[k][c]) for c in cols)))
tot = collections.Counter()
for k in rows:
    tot.update(rows[k])
print("TOTAL | %d | %s" % (tot["n"], " | ".join(str(tot[c]) for c in cols)))
