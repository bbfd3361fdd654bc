#!/usr/bin/env python
"""Dynamic SASS summary of one kernel from `ncu -i rep --page source --csv`: executed warp-instructions
per opcode, stall samples per opcode, local-memory traffic.  usage: ncu_sass_summary.py report.ncu-rep [iters]"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
warp_iters = float(sys.argv[2]) if len(sys.argv) > 2 else None  # total warp-iterations, to normalise
import os

# NCU_IMPORT_ARGS="--launch-skip 1 --launch-count 1" picks a launch of the report other than the first
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + os.environ.get("NCU_IMPORT_ARGS", "").split(),
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
ex = collections.Counter()
st = collections.Counter()
thr = collections.Counter()
tot = 0
for r in rows[2:]:
    if r and r[0] == "Address":  # the report holds further views / launches: the first SASS view is enough
        break
    if len(r) < len(hdr):
        continue
    m = re.match(r"\s*(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", r[ci["Source"]])
    if not m:
        continue
    op = m.group(1)
    n = int(r[ci["Instructions Executed"]] or 0)
    ex[op] += n
    thr[op] += int(r[ci["Thread Instructions Executed"]] or 0)
    st[op] += int(r[ci["Warp Stall Sampling (All Samples)"]] or 0)
    tot += n
fp64 = sum(ex[k] for k in ("DFMA", "DMUL", "DADD", "DSETP", "MUFU"))
print(f"warp-instructions executed: {tot:,}   FP64-pipe: {fp64:,} ({100*fp64/tot:.1f}%)")
if warp_iters:
    print(f"per warp-iteration: total {tot/warp_iters:.0f}, FP64 {fp64/warp_iters:.0f}")
stot = sum(st.values())
for op, n in ex.most_common(24):
    print(f"  {op:10s} {n:>14,} {100*n/tot:5.1f}%  lanes {thr[op]/max(n,1):5.1f}  stall-samples {100*st[op]/max(stot,1):5.1f}%")

# where the not-issued stall samples fall, by reason
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
rs = collections.Counter()
for r in rows[2:]:
    if r and r[0] == "Address":
        break
    if len(r) < len(hdr):
        continue
    for k in reasons:
        try:
            rs[k] += int(r[ci[k]] or 0)
        except ValueError:
            pass
tot_s = sum(rs.values())
print("stall samples by reason: " + ", ".join(f"{k[6:]} {100*v/max(tot_s,1):.1f}%" for k, v in rs.most_common(10)))
