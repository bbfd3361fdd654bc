"""Diagnostic: in-order issue model of ONE warp running alone through a straight-line SASS path.

Blackwell's FP64 instructions carry no compiler stall counts for their result latency (the hardware
interlocks), so the single-warp time of one Newton trip — which is what the launch tail costs — has to be
estimated from the register dependencies.  Model: in-order issue, one instruction per clock, an FP64
instruction occupies its pipe for 2 clocks, a result is usable LAT[pipe] clocks after issue.

usage: python tools/sass_sim.py file.sass start:end [start:end ...] [--lat-fp64 N] [-v]
Slow-path regions (a forward predicated BRA that jumps over a CALL) are skipped.
"""
import re
import sys

LAT = {"fp64": 8, "alu": 4, "mufu": 18, "ldc": 12, "lds": 24, "ldg": 400, "uni": 6}
args = [a for a in sys.argv[1:] if not a.startswith("-")]
if "--lat-fp64" in sys.argv:
    LAT["fp64"] = int(sys.argv[sys.argv.index("--lat-fp64") + 1])
    args = [a for a in args if a != str(LAT["fp64"])]
verbose = "-v" in sys.argv

ins = []
for ln in open(args[0]):
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
addr_index = {a: i for i, (a, _) in enumerate(ins)}


def pipe_of(op):
    b = op.split(".")[0]
    if b in ("DFMA", "DMUL", "DADD", "DSETP"):
        return "fp64"
    if b == "MUFU":
        return "mufu"
    if b in ("LDC", "LDCU"):
        return "ldc"
    if b in ("LDS", "LDL"):
        return "lds"
    if b in ("LDG", "LD", "ATOMG", "ATOM"):
        return "ldg"
    if b.startswith("U") and b not in ("UNPACK",):
        return "uni"
    return "alu"


def regs_of(tok, wide):
    """registers named by one operand token -> list of names (pairs expanded for 64-bit operands)"""
    out = []
    for m in re.finditer(r"\b(UR|R|UP|P)(\d+)\b", tok):
        kind, num = m.group(1), int(m.group(2))
        if kind in ("R", "UR"):
            out.append(f"{kind}{num}")
            if wide >= 2:
                out.append(f"{kind}{num + 1}")
            if wide >= 4:
                out.append(f"{kind}{num + 2}")
                out.append(f"{kind}{num + 3}")
        else:
            out.append(f"{kind}{num}")
    return out


def decode(text):
    guard = []
    m = re.match(r"@(!?)(U?P\d+)\s+(.*)", text)
    if m:
        guard = [m.group(2)]
        text = m.group(3)
    parts = text.split(None, 1)
    op = parts[0]
    ops = [t.strip() for t in parts[1].split(",")] if len(parts) > 1 else []
    base = op.split(".")[0]
    wide64 = base in ("DFMA", "DMUL", "DADD", "DSETP") or ".64" in op
    wide = 4 if ".128" in op else (2 if wide64 else 1)
    dst, src = [], list(guard)
    if base in ("STG", "STL", "STS", "ST", "BRA", "EXIT", "BSSY", "BSYNC", "CALL", "RET", "NOP", "WARPSYNC", "RED"):
        for t in ops:
            src += regs_of(t, wide if base.startswith("ST") else 1)
        return op, dst, src
    ndst = 1
    if base in ("DSETP", "ISETP", "FSETP", "PLOP3", "UISETP") or (base == "LOP3" and ops and ops[0].startswith("P")):
        ndst = 2
    for k, t in enumerate(ops):
        if k < ndst:
            w = wide if not (base == "DSETP") else 1
            if base in ("MUFU",):
                w = 1
            dst += regs_of(t, w)
        else:
            w = wide
            if base == "MUFU":
                w = 1
            if base in ("LDG", "LDL", "LDS", "LDC", "LDCU", "LD"):
                w = 2 if "R" in t and "[" in t and base in ("LDG", "LD") else 1
            src += regs_of(t, w)
    return op, dst, src


def path(ranges):
    seq = []
    for r in ranges:
        a, b = (int(v, 16) for v in r.split(":"))
        i = addr_index[a]
        while i < len(ins) and ins[i][0] <= b:
            addr, text = ins[i]
            m = re.match(r"@!?U?P\d+\s+BRA\s+(?:P\d+,\s*)?0x([0-9a-f]+)", text)
            if m:
                tgt = int(m.group(1), 16)
                if tgt > addr and tgt in addr_index and any("CALL" in ins[j][1] for j in range(i, addr_index[tgt])):
                    seq.append((addr, text))
                    i = addr_index[tgt]
                    continue
            seq.append((addr, text))
            i += 1
    return seq


seq = path(args[1:])
ready = {}
t = 0
fp64_free = 0
n_fp64 = 0
stall_by = {}
for addr, text in seq:
    op, dst, src = decode(text)
    p = pipe_of(op)
    t0 = t
    need = max([ready.get(r, 0) for r in src] + [0])
    issue = max(t, need)
    why = "dep" if need > t else ""
    if p == "fp64":
        n_fp64 += 1
        if fp64_free > issue:
            issue = fp64_free
            why = "pipe"
        fp64_free = issue + 2
    if issue > t0:
        stall_by[why] = stall_by.get(why, 0) + (issue - t0)
    for r in dst:
        if r not in ("RZ", "PT", "URZ", "UPT"):
            ready[r] = issue + LAT[p]
    if verbose:
        print(f"{addr:05x} t={issue:6d} (+{issue - t0:3d} {why:4s}) {text}")
    t = issue + 1
print(f"{len(seq)} instructions ({n_fp64} FP64), {t} clocks for one warp alone = {t / len(seq):.2f} clk/instr; "
      f"FP64 pipe floor {2 * n_fp64}; stall clocks {stall_by}; LAT={LAT}")
