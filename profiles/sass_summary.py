#!/usr/bin/env python3
"""SASS evidence for the built library: per kernel, the instruction count and the mnemonics that show how it talks to memory
(bulk async copies UBLKCP, mbarrier SYNCS, reductions REDG, shared-memory atomics ATOMS, 128/256-bit LDG/STG, PRMT, ...).
  python profiles/sass_summary.py [kwage_b200/lib/libkwage_cuda.so] > profiles/rNN_sass_summary.md"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "kwage_b200/lib/libkwage_cuda.so"
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
archs = sorted(set(re.findall(r"arch = (sm_\w+)", txt)))
KEYS = ["UBLKCP", "SYNCS", "REDG", "ATOMS", "ATOMG", "LDG.256", "LDG.128", "LDG.64", "STG.128", "STG.64", "LDS.128", "STS.128",
        "PRMT", "SHFL", "VOTE", "BAR", "LDGSTS", "POPC", "IMAD", "LOP3", "SHF"]
cur, per = None, collections.OrderedDict()
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0].replace("void ", "").replace("kwg::", "")
        per[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        per[cur]["_total"] += 1
        parts = op.split(".")
        for k in KEYS:
            kp = k.split(".")
            if parts[0] == kp[0] and all(x in parts[1:] for x in kp[1:]):
                per[cur][k] += 1
print("# SASS summary of %s (cuobjdump -sass), cubins for: %s\n" % (lib, ", ".join(archs)))
print("Static instruction counts per kernel; only mnemonics that occur are listed.\n")
print("| kernel | instructions | memory / sync mnemonics |")
print("|---|---|---|")
for k, c in per.items():
    items = ["%s x%d" % (m, c[m]) for m in KEYS if c[m] and m not in ("IMAD", "LOP3", "SHF")]
    print("| `%s` | %d | %s |" % (k, c["_total"], ", ".join(items)))
tot = collections.Counter()
for c in per.values():
    tot.update(c)
print("\nwhole library: %d instructions; UBLKCP (bulk async copy) %d, SYNCS (mbarrier) %d, REDG %d, ATOMS %d, PRMT %d; no HMMA/UTCMMA/tcgen05: the path "
      "has no dense contraction (BASELINE.json north_star)." % (tot["_total"], tot["UBLKCP"], tot["SYNCS"], tot["REDG"], tot["ATOMS"], tot["PRMT"]))
