#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of an sm_100a shared object (cuobjdump -sass), to check the
FP64 instruction mix before spending GPU time.  usage: sass_hist.py lib.so [name-substring]"""
import collections
import re
import subprocess
import sys

so = sys.argv[1]
pat = sys.argv[2] if len(sys.argv) > 2 else ""
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
fn = None
hist = collections.defaultdict(collections.Counter)
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", line)
    if m and fn:
        hist[fn][m.group(1)] += 1
for f, h in hist.items():
    if pat not in f:
        continue
    tot = sum(h.values())
    fp64 = sum(v for k, v in h.items() if k in ("DFMA", "DMUL", "DADD", "DSETP", "MUFU"))
    print(f"== {f}\n   total {tot}  fp64-pipe {fp64} ({100.0*fp64/tot:.1f}%)")
    print("   " + "  ".join(f"{k}:{v}" for k, v in h.most_common(28)))
