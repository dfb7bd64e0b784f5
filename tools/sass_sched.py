"""Decode the scheduling control bits of a kernel's SASS (development aid, no GPU needed).
Volta+ 128-bit encoding: bits 105-108 stall count, 109 yield, 110-112 write barrier, 113-115 read barrier, 116-121 wait mask, 122-125 reuse.
usage: python tools/sass_sched.py <object-or-so> <kernel-substring> [--loop] [--dump]"""
import re, subprocess, sys
from collections import Counter

def parse(path, pat):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    cur = None; ins = {}; pend = None
    for line in out.splitlines():
        m = re.match(r"\s+Function : (\S+)", line)
        if m: cur = m.group(1); ins[cur] = []; continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\* 0x([0-9a-f]{16}) \*/", line)
        if m and cur:
            pend = [int(m.group(1), 16), m.group(2).strip(), int(m.group(3), 16)]; continue
        m = re.match(r"\s+/\* 0x([0-9a-f]{16}) \*/", line)
        if m and pend:
            hi = int(m.group(1), 16); w = pend[2] | (hi << 64)
            ctl = (w >> 105) & ((1 << 21) - 1)
            pend.append({"stall": ctl & 15, "yield": (ctl >> 4) & 1, "wbar": (ctl >> 5) & 7, "rbar": (ctl >> 8) & 7, "wait": (ctl >> 11) & 63, "reuse": (ctl >> 17) & 15})
            ins[cur].append(pend); pend = None
    return {k: v for k, v in ins.items() if pat in k}

def hot_loop(ins):
    best = 0; lo, hi = 0, 1 << 60
    for addr, txt, _, c in ins:
        m = re.match(r"@!?U?P\d+\s+BRA.*0x([0-9a-f]+)\s*$", txt)
        if m:
            t = int(m.group(1), 16)
            if t < addr and addr - t > best: best = addr - t; lo, hi = t, addr
    return [i for i in ins if lo <= i[0] <= hi]

if __name__ == "__main__":
    path, pat = sys.argv[1], sys.argv[2]
    for name, ins in parse(path, pat).items():
        body = hot_loop(ins) if "--loop" in sys.argv else ins
        st = sum(i[3]["stall"] for i in body)
        ops = Counter(re.sub(r"^@!?U?P\d+\s+", "", i[1]).split()[0] for i in body)
        print(f"{name[:80]}: {len(body)} instr, sum of stall counts {st}, mean {st/len(body):.2f}")
        print("  ", ", ".join(f"{k}:{v}" for k, v in ops.most_common(40)))
        if "--dump" in sys.argv:
            for addr, txt, _, c in body:
                print(f"{addr:05x} s{c['stall']:<2d} {'Y' if c['yield'] else ' '} w{c['wbar'] if c['wbar']!=7 else '-'} r{c['rbar'] if c['rbar']!=7 else '-'} m{c['wait']:02x} u{c['reuse']:x}  {txt}")
