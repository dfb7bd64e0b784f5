"""Shared-memory bank-conflict model of the NTT tile exchange (development aid, no GPU needed).

For a (word bits, log2 N) configuration it replays the tile addresses every warp touches in every pass (ntt_plan.h shapes,
ntt_engine.cuh tile_load / tile_store / tile_to_gmem layouts) and counts wavefronts per warp request against the minimum
(32 banks x 4 bytes; 64-bit accesses are served per half-warp, 128-bit accesses per quarter-warp).
  python tools/bank_conflicts.py            # all shapes the library instantiates
"""
import sys


def plan(n, wb):
    emax = (4 if n == 10 else 5) if wb == 64 else (5 if n >= 12 else 6)
    npass = (n + emax - 1) // emax
    e = (n + npass - 1) // npass
    r = [n - e * (npass - 1)] + [e] * (npass - 1)
    s0 = [0] + [r[0] + (i - 1) * e for i in range(1, npass)]
    hi = [n - s for s in s0]
    c = [h - e for h in hi]
    return e, npass, r, s0, hi, c


def addr_pad(pos, e, padw):
    return pos + (pos >> e) * padw


def addr_swz(pos):
    return pos ^ (((pos >> 5) & 3) << 2) ^ (((pos >> 8) & 1) << 4)


def wavefronts(word_addrs, words_per_access):
    """word_addrs: per lane, first 4-byte word address of the access; returns (wavefronts, minimum)"""
    lanes_per_phase = 32 // words_per_access
    total = 0
    for ph in range(0, 32, lanes_per_phase):
        banks = {}
        for a in word_addrs[ph:ph + lanes_per_phase]:
            for w in range(words_per_access):
                banks.setdefault((a + w) & 31, set()).add(a + w)
        total += max(len(v) for v in banks.values())
    return total, words_per_access


def analyse(wb, n, swz=False, adj=True):
    e, npass, r, s0, hi, c = plan(n, wb)
    E = 1 << e
    wsz = wb // 32                     # 4-byte words per element
    vec = 16 // (wb // 8)              # elements per 16-byte vector
    padw = vec
    tpu = (1 << n) >> e
    A = (lambda pos: addr_swz(pos)) if swz else (lambda pos: addr_pad(pos, e, padw))
    out = []
    for i in range(npass):
        if npass == 1:
            break
        tot = mn = 0
        for w0 in range(0, tpu, 32):
            lanes = [w0 + l for l in range(min(32, tpu))]
            cb = e - r[0] if (wb == 64 and i == 0 and adj) else 0
            cb = cb if (cb == 1 and n == 14) else 0  # NttCfg::CB (the shipped default)
            if cb >= 1:  # pass 0 with adjacent columns (NttCfg::ADJ): 2^cb consecutive words per butterfly index, 16-byte vectors
                for kh in range(1 << r[0]):
                    for v in range((1 << cb) // vec):
                        addrs = [A((kh << (n - r[0])) | (t << cb) | (v * vec)) * wsz for t in lanes]
                        wf, m = wavefronts(addrs + [addrs[0]] * (32 - len(addrs)), 4)
                        tot += wf; mn += m
            elif c[i] == 0:  # row per thread, 16-byte vectors
                for v in range(E // vec):
                    addrs = [A(t * E + v * vec) * wsz for t in lanes]
                    wf, m = wavefronts(addrs + [addrs[0]] * (32 - len(addrs)), 4)
                    tot += wf; mn += m
            else:
                for k in range(E):
                    addrs = [A(((t >> c[i]) << hi[i]) | (k << c[i]) | (t & ((1 << c[i]) - 1))) * wsz for t in lanes]
                    wf, m = wavefronts(addrs + [addrs[0]] * (32 - len(addrs)), wsz)
                    tot += wf; mn += m
        out.append(("pass %d (r=%d, c=%d)" % (i, r[i], c[i]), tot, mn))
    # coalesced 16-byte-per-lane copy between tile and global memory
    if npass > 1:
        tot = mn = 0
        chunks = (1 << n) // vec
        for ch0 in range(0, chunks, 32):
            addrs = [A((ch0 + l) * vec) * wsz for l in range(32)]
            wf, m = wavefronts(addrs, 4)
            tot += wf; mn += m
        out.append(("copy (16 B / lane)", tot, mn))
    return e, r, out


if __name__ == "__main__":
    shapes = [(64, n) for n in range(6, 15)] + [(32, n) for n in range(7, 16)]
    for wb, n in shapes:
        e, r, res = analyse(wb, n)
        if not res:
            continue
        line = "  ".join("%s: %d/%d" % (nm, t, m) for nm, t, m in res)
        print("u%d N=2^%-2d shape %s  padded rows    | %s" % (wb, n, tuple(r), line))
        if wb == 32 and n == 12:
            e, r, res = analyse(wb, n, swz=True)
            print("u%d N=2^%-2d shape %s  XOR swizzle    | %s" % (wb, n, tuple(r), "  ".join("%s: %d/%d" % (nm, t, m) for nm, t, m in res)))
