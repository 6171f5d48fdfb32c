#!/usr/bin/env python3
"""Opcode census of the shipped library: per kernel, how many tcgen05 / TMEM / TMA instructions it contains.

    python tools/sass_census.py [aicity_action_b200/lib/libmvit_b200.so] > profiles/sass_census_r2.txt
UTCHMMA = tcgen05.mma (bf16), UTCBAR = tcgen05.commit, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA tensor load /
store, SYNCS = mbarrier, FFMA2 / FADD2 / FMUL2 = packed fp32x2 arithmetic, MUFU = special-function unit."""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "aicity_action_b200/lib/libmvit_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
KEYS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMACCTL", "SYNCS", "LDGSTS", "FFMA2", "FADD2", "FMUL2",
        "MUFU", "HMMA", "USETMAXREG"]
per = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        per[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        op = m.group(1)
        per[cur]["_total"] += 1
        for k in KEYS:
            if op.startswith(k):
                per[cur][k] += 1
print(f"{'kernel':70s} {'instrs':>7s} " + " ".join(f"{k:>8s}" for k in KEYS))
tot = collections.Counter()
for name, c in per.items():
    if not any(c[k] for k in KEYS[:8]) and c["FFMA2"] == 0:
        continue
    print(f"{name[-70:]:70s} {c['_total']:7d} " + " ".join(f"{c[k]:8d}" for k in KEYS))
    tot.update(c)
print(f"{'TOTAL (listed kernels)':70s} {tot['_total']:7d} " + " ".join(f"{tot[k]:8d}" for k in KEYS))
