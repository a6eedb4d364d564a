#!/usr/bin/env python
"""Opcode histogram per kernel of a built library: the SASS evidence for TMA (UTMALDG/UTMASTG), cp.async (LDGSTS), mbarrier
(SYNCS), fp64 (DFMA/DADD/DMUL) and shared-memory traffic (LDS/STS).    python tools/sass_hist.py cales_b200/libcales_b200.so"""
import collections
import re
import subprocess
import sys

KEEP = ("UTMALDG", "UTMASTG", "UTMACMDFLUSH", "LDGSTS", "SYNCS", "DFMA", "DADD", "DMUL", "MUFU", "LDS", "STS", "LDG", "STG", "BAR", "SHFL", "FENCE", "ATOMG", "REDG")
lib = sys.argv[1] if len(sys.argv) > 1 else "cales_b200/libcales_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
fn, hist, tot = None, collections.OrderedDict(), collections.Counter()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        hist[fn] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and fn:
        tot[fn] += 1
        if m.group(2) in KEEP:
            hist[fn][m.group(2)] += 1
names = subprocess.run(["c++filt"], input="\n".join(hist), capture_output=True, text=True).stdout.splitlines()
for raw, nm in sorted(zip(hist, names), key=lambda t: t[1]):
    short = re.sub(r"\(.*", "", nm)
    print("%-70s total %6d  %s" % (short[:70], tot[raw], " ".join("%s=%d" % (k, hist[raw][k]) for k in KEEP if hist[raw][k])))
