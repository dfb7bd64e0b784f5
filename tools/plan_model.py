"""Executable model (pure Python ints) of the multi-pass register-radix plan in nfllib_b200/csrc/ntt_plan.h.
Development aid: checks the index algebra (window k, group g, twiddle slot e_idx) against the CPU oracle
before it is written as CUDA.  Not used by the product or the bench."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np
from oracle_lib import Oracle, golden_params, random_polys


def plan(n, wb):
    emax = (4 if n == 10 else 5) if wb == 64 else (5 if n >= 12 else 6)  # mirrors plan_emax() in ntt_plan.h
    npass = (n + emax - 1) // emax
    e = (n + npass - 1) // npass
    r = [n - e * (npass - 1)] + [e] * (npass - 1)
    s0 = [0]
    for i in range(1, npass):
        s0.append(s0[-1] + r[i - 1])
    return e, r, s0


def brv(v, bits):
    return int(format(v, f"0{bits}b")[::-1], 2) if bits else 0


def tables(n, wb, p, root, kmax):
    N = 1 << n
    psi = root
    k = kmax
    while k > N:
        psi = psi * psi % p
        k >>= 1
    ipsi = pow(psi, p - 2, p)
    ninv = pow(N, p - 2, p)
    e, r, s0 = plan(n, wb)
    fw = [0] * N
    iw = [0] * N
    for i in range(len(r)):
        G = 1 << s0[i]
        off = (1 << s0[i]) - 1
        for q in range(r[i]):
            for kk in range(1 << q):
                for g in range(G):
                    eidx = (1 << q) - 1 + kk
                    kidx = (1 << (s0[i] + q)) + (g << q) + kk
                    ex = brv(kidx, n)
                    fw[off + eidx * G + g] = pow(psi, ex, p)
                    v = pow(ipsi, ex, p)
                    if kidx == 1:
                        v = v * ninv % p
                    iw[off + eidx * G + g] = v
    iw[N - 1] = ninv
    return fw, iw


def fwd(x, n, wb, p, fw):
    N = 1 << n
    e, r, s0 = plan(n, wb)
    E = 1 << e
    x = list(x)
    for i in range(len(r)):
        hi = n - s0[i]
        c = hi - e
        G = 1 << s0[i]
        off = (1 << s0[i]) - 1
        for tid in range(N >> e):
            g, l = tid >> c, tid & ((1 << c) - 1)
            pos = [(g << hi) | (k << c) | l for k in range(E)]
            v = [x[q_] for q_ in pos]
            for q in range(r[i]):
                bit = e - 1 - q
                for k in range(E):
                    if k & (1 << bit):
                        continue
                    eidx = (1 << q) - 1 + (k >> (e - q))
                    w = fw[off + eidx * G + g]
                    X, Y = v[k], v[k | (1 << bit)]
                    T = Y * w % p
                    v[k], v[k | (1 << bit)] = (X + T) % p, (X - T) % p
            for k in range(E):
                x[pos[k]] = v[k]
    return x


def inv(x, n, wb, p, iw):
    N = 1 << n
    e, r, s0 = plan(n, wb)
    E = 1 << e
    x = list(x)
    ninv = iw[N - 1]
    for i in reversed(range(len(r))):
        hi = n - s0[i]
        c = hi - e
        G = 1 << s0[i]
        off = (1 << s0[i]) - 1
        for tid in range(N >> e):
            g, l = tid >> c, tid & ((1 << c) - 1)
            pos = [(g << hi) | (k << c) | l for k in range(E)]
            v = [x[q_] for q_ in pos]
            for q in reversed(range(r[i])):
                bit = e - 1 - q
                last = (i == 0 and q == 0)
                for k in range(E):
                    if k & (1 << bit):
                        continue
                    eidx = (1 << q) - 1 + (k >> (e - q))
                    w = iw[off + eidx * G + g]
                    U, V = v[k], v[k | (1 << bit)]
                    v[k] = (U + V) * (ninv if last else 1) % p
                    v[k | (1 << bit)] = (U - V) * w % p
            for k in range(E):
                x[pos[k]] = v[k]
    return x


if __name__ == "__main__":
    for bits, wb, n in [(64, 64, 2), (64, 64, 3), (64, 64, 6), (64, 64, 7), (64, 64, 10), (64, 64, 11), (32, 32, 7), (32, 32, 12), (32, 32, 13), (16, 32, 9)]:
        N = 1 << n
        g = golden_params(bits)
        p, root, kmax = g["P"][0], g["roots"][0], g["kmax"]
        fw, iw = tables(n, wb, p, root, kmax)
        a = random_polys(bits, N, 1, 1, 3 + n)
        o = Oracle(bits, N, 1)
        ref = o.run("fwd", a)[0, 0]
        mine = fwd([int(v) for v in a[0, 0]], n, wb, p, fw)
        ok1 = [int(v) for v in ref] == mine
        back = inv(mine, n, wb, p, iw)
        ok2 = back == [int(v) for v in a[0, 0]]
        refinv = o.run("inv", a)[0, 0]
        ok3 = [int(v) for v in refinv] == inv([int(v) for v in a[0, 0]], n, wb, p, iw)
        print(bits, n, plan(n, wb), ok1, ok2, ok3)
        assert ok1 and ok2 and ok3
