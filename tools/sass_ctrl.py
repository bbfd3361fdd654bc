"""Diagnostic: decode the scheduling control bits of a cuobjdump -sass listing (stall count, yield, write/read
barrier, wait mask) and print a per-range issue-cycle estimate for ONE warp running alone.
usage: python tools/sass_ctrl.py file.sass [start_addr end_addr]"""
import re
import sys

ins = []
cur = None
for ln in open(sys.argv[1]):
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);\s+/\* 0x([0-9a-f]{16}) \*/", ln)
    if m:
        cur = {"addr": int(m.group(1), 16), "text": m.group(2).strip(), "lo": int(m.group(3), 16)}
        continue
    m = re.match(r"\s+/\* 0x([0-9a-f]{16}) \*/", ln)
    if m and cur is not None:
        hi = int(m.group(1), 16)
        cur["stall"] = (hi >> 41) & 0xF
        cur["yield"] = (hi >> 45) & 1
        cur["wrbar"] = (hi >> 46) & 7
        cur["rdbar"] = (hi >> 49) & 7
        cur["wait"] = (hi >> 52) & 0x3F
        ins.append(cur)
        cur = None
lo = int(sys.argv[2], 16) if len(sys.argv) > 2 else 0
hi_ = int(sys.argv[3], 16) if len(sys.argv) > 3 else 1 << 30
sel = [i for i in ins if lo <= i["addr"] < hi_]
tot = sum(i["stall"] for i in sel)
print(f"{len(sel)} instructions, sum of stall counts {tot}, mean {tot/len(sel):.2f}")
from collections import Counter
opst = Counter()
opn = Counter()
wb = Counter()
for i in sel:
    op = i["text"].split()[0]
    if op.startswith("@"):
        op = i["text"].split()[1]
    op = op.split(".")[0]
    opst[op] += i["stall"]
    opn[op] += 1
    if i["wrbar"] != 7:
        wb[op] += 1
for op, n in opn.most_common(14):
    print(f"  {op:8s} n={n:5d} stall-sum={opst[op]:6d} mean={opst[op]/n:5.2f}  sets-write-barrier={wb[op]}")
if "-v" in sys.argv:
    for i in sel:
        print(f"{i['addr']:05x} st={i['stall']:2d} y={i['yield']} wb={i['wrbar']} rb={i['rdbar']} wait={i['wait']:06b}  {i['text']}")
