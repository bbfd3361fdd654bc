#!/usr/bin/env python
"""Diagnostic: where a kernel's executed instructions and stall samples fall, by source function.
usage: ncu_regions.py report.ncu-rep annotated.sass('nvdisasm -gi' text) kernel-mangled-name"""
import collections
import csv
import io
import re
import subprocess
import sys

rep, dis, kname = sys.argv[1:4]
# ---- address -> innermost (file, line) from nvdisasm -gi
addr_src = {}
cur = None
fresh = True
inside = False
for ln in open(dis):
    if ln.startswith(".text."):
        inside = ln.strip() == f".text.{kname}:"
        continue
    if not inside:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        f = m.group(1).split("/")[-1]
        if fresh and f in ("ccp_core.h", "ccp_project.cu", "ccp_device.cuh"):  # innermost frame in our own sources
            cur = (f, int(m.group(2)))
            fresh = False
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", ln)
    if m:
        addr_src[int(m.group(1), 16)] = cur
        fresh = True
# ---- line -> function for the files of interest
def func_map(path):
    out = []
    name = None
    for i, ln in enumerate(open(path), 1):
        m = re.match(r"\s*(?:template.*>\s*)?(?:CCP_HD|__device__ __forceinline__|static inline)\s+[\w:<>\*& ]*?(\w+)\(", ln)
        if m:
            name = m.group(1)
        if ln.startswith("ccp_project_kernel("):
            name = "kernel loop / epilogue"
        out.append(name)
    return out
import os
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
fm = {f: func_map(os.path.join(root, "closed_chain_motion_planner_b200", "csrc", f)) for f in ("ccp_core.h", "ccp_project.cu", "ccp_device.cuh")}
import os

# NCU_IMPORT_ARGS="--launch-skip 1 --launch-count 1" picks a launch of the report other than the first
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + os.environ.get("NCU_IMPORT_ARGS", "").split(),
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = collections.defaultdict(lambda: collections.Counter())
base = int(rows[2][ci["Address"]], 16)  # ncu prints absolute addresses; the listing is function-relative
for r in rows[2:]:
    if r and r[0] == "Address":
        break
    if len(r) < len(hdr):
        continue
    try:
        a = int(r[ci["Address"]], 16) - base
    except ValueError:
        continue
    src = addr_src.get(a)
    reg = "?"
    if src:
        f, l = src
        if f in fm and l - 1 < len(fm[f]):
            reg = f"{fm[f][l - 1]}"
        else:
            reg = f
    A = agg[reg]
    A["inst"] += int(r[ci["Instructions Executed"]] or 0)
    A["samples"] += int(r[ci["Warp Stall Sampling (All Samples)"]] or 0)
    for k in reasons:
        A[k] += int(r[ci[k]] or 0)
ti = sum(v["inst"] for v in agg.values())
ts = sum(v["samples"] for v in agg.values())
print(f"{'region':28s} {'inst%':>6s} {'samples%':>8s}  samples/inst(rel)  top stall reasons")
for reg, A in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:24]:
    top = sorted(((A[k], k[6:]) for k in reasons), reverse=True)[:4]
    rel = (A["samples"] / ts) / max(A["inst"] / ti, 1e-9)
    print(f"{reg:28s} {100*A['inst']/ti:6.1f} {100*A['samples']/ts:8.1f}  {rel:6.2f}   " + ", ".join(f"{n} {100*v/max(A['samples'],1):.0f}%" for v, n in top))
