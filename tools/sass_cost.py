"""Static cost model of a kernel from its SASS (development aid, no GPU needed).
Integer code on sm_100 is issue-bound (profiles/r01_integer_pipe_model.md): ~2 cycles per IMAD / IMAD.WIDE / IADD3-class warp
instruction and sub-partition, IMAD.HI about twice that; `issue_cyc` (the instruction count of the hot loop) is the figure that
tracks measured time.  The fmaheavy / alu columns reproduce ncu's pipe-busy counters (IMAD.WIDE = 4 busy cycles, IMAD = 2, ALU = 2).
usage: python tools/sass_cost.py <object-or-so> <kernel-name-substring> [butterflies]"""
import re
import subprocess
import sys
from collections import Counter

ALU = ("IADD3", "LOP3", "SEL", "ISETP", "SHF", "LEA", "MOV", "PRMT", "VIADD", "VIADDMNMX", "IMNMX", "VIMNMX", "PLOP3", "FSEL", "IABS", "BMSK", "SGXT", "FLO", "POPC", "CS2R")


def kernels(path, hot_loop=True):
    """opcode list per kernel; with hot_loop, only the instructions inside the widest backward branch (the unit loop)."""
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    cur, body = None, {}
    for line in out.splitlines():
        m = re.match(r"\s+Function : (\S+)", line)
        if m:
            cur = m.group(1)
            body[cur] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(@!?U?P\w+\s+)?([A-Z0-9_.]+)\s*([^;]*);", line)
        if m and cur:
            body[cur].append((int(m.group(1), 16), m.group(3), m.group(4), bool(m.group(2))))
    res = {}
    for name, ins in body.items():
        lo, hi = 0, 1 << 60
        if hot_loop:
            best = 0
            for addr, op, args, pred in ins:
                if op.startswith("BRA") and pred:  # the unit loop closes with a predicated backward branch (an unpredicated one is the mbarrier wait)
                    t = re.search(r"0x([0-9a-f]+)\s*$", args.strip())
                    if t and int(t.group(1), 16) < addr and addr - int(t.group(1), 16) > best:
                        best = addr - int(t.group(1), 16)
                        lo, hi = int(t.group(1), 16), addr
        res[name] = [op for addr, op, _, _ in ins if lo <= addr <= hi]
    return res


def cost(ops):
    c = Counter(ops)
    wide = sum(v for k, v in c.items() if k.startswith("IMAD.WIDE"))
    hi = sum(v for k, v in c.items() if k.startswith("IMAD.HI"))
    imad = sum(v for k, v in c.items() if k.startswith("IMAD") and not k.startswith("IMAD.WIDE") and not k.startswith("IMAD.HI"))
    alu = sum(v for k, v in c.items() if k.split(".")[0] in ALU)
    lsu = sum(v for k, v in c.items() if k.split(".")[0] in ("LDS", "STS", "LDG", "STG", "LD", "ST", "LDSM"))
    n = len(ops)
    return {"n": n, "wide": wide, "imad": imad, "alu": alu, "lsu": lsu, "hi": hi, "fmaheavy_cyc": 4 * wide + 6 * hi + 2 * imad, "alu_cyc": 2 * alu, "issue_cyc": n}


if __name__ == "__main__":
    path, pat = sys.argv[1], sys.argv[2]
    nb = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    for name, ops in kernels(path).items():
        if pat in name:
            r = cost(ops)
            line = f"{name[:70]:70s} n={r['n']:5d} wide={r['wide']:4d} hi={r['hi']:4d} imad={r['imad']:4d} alu={r['alu']:4d} lsu={r['lsu']:3d} | fmaheavy={r['fmaheavy_cyc']:5d} alu={r['alu_cyc']:5d} issue={r['issue_cyc']:5d}"
            if nb:
                line += f" | per butterfly: fma={r['fmaheavy_cyc'] / nb:5.1f} alu={r['alu_cyc'] / nb:5.1f} issue={r['issue_cyc'] / nb:5.1f}"
            print(line)
