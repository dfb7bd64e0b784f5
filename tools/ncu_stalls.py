"""Per-opcode warp-stall profile of a kernel from an ncu report taken with --import-source on (runs where ncu is installed, no GPU
needed): for every SASS opcode class, executed warp instructions, share of all stall samples, and its top stall reasons.
usage: python tools/ncu_stalls.py report.ncu-rep [kernel-index]"""
import collections
import csv
import io
import re
import subprocess
import sys


def main(path, which=None):
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    for n, i0 in enumerate(starts):
        if which is not None and n != which:
            continue
        hdr = rows[i0 + 1]
        end = starts[n + 1] if n + 1 < len(starts) else len(rows)
        col = {h: i for i, h in enumerate(hdr)}
        reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        agg, cnt = collections.defaultdict(collections.Counter), collections.Counter()
        for r in rows[i0 + 2:end]:
            if len(r) < len(hdr):
                continue
            src = re.sub(r"^@!?U?P\d+\s+", "", r[col["Source"]].strip())
            if not src:
                continue
            op = src.split()[0]
            key = op.split(".")[0] + (".WIDE" if "WIDE" in op else "") + (".X" if op.endswith(".X") else "")
            cnt[key] += int(r[col["Instructions Executed"]])
            for h in reasons:
                agg[key][h] += int(r[col[h]])
        tot = sum(sum(v.values()) for v in agg.values()) or 1
        print(f"== {rows[i0][1]}: {sum(cnt.values())} warp instructions, {tot} stall samples")
        overall = collections.Counter()
        for v in agg.values():
            overall.update(v)
        print("   all opcodes: " + ", ".join(f"{h[6:]} {100 * c / tot:.1f}%" for h, c in overall.most_common(9)))
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1].values()))[:14]:
            s = sum(v.values()) or 1
            print(f"   {k:12s} executed {cnt[k]:10d}  samples {100 * s / tot:5.1f}%  | " + ", ".join(f"{h[6:]} {100 * c / s:.0f}%" for h, c in v.most_common(5)))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else None)
