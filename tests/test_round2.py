"""Round-2 additions: SURVEY Appendix C known answers (the one vector set not produced by this repo's scripts), the device
`==` / `!=`, compile-time expression programs, the peer-memory residue gather, the pooled scratch and the BASELINE.json
configurations at their FULL batch sizes.  GPU tests go through the C ABI; comparisons are bit for bit."""
import hashlib
import json
import multiprocessing as mp
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle_lib import GOLDEN, DTYPES, Oracle, Ref, have_ref, golden_params, random_polys

import nfllib_b200 as nb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


# ---- SURVEY.md Appendix C ----------------------------------------------------------------------------------------------

def appendix_c_cases():
    with open(os.path.join(GOLDEN, "appendix_c.json")) as f:
        return json.load(f)["cases"]


def appendix_c_inputs(bits, N, M):
    P = golden_params(bits)["P"]
    i = np.arange(N, dtype=np.uint64)
    A = np.empty((1, M, N), dtype=DTYPES[bits])
    B = np.empty((1, M, N), dtype=DTYPES[bits])
    for cm in range(M):
        A[0, cm] = ((i * np.uint64(2654435761) + np.uint64(cm * 40503 + 1)) % np.uint64(P[cm])).astype(DTYPES[bits])
        B[0, cm] = ((i * np.uint64(2246822519) + np.uint64(cm * 40503 + 7)) % np.uint64(P[cm])).astype(DTYPES[bits])
    return A, B


CASE_IDS = [f"u{c['bits']}_n{c['N']}_m{c['M']}" for c in appendix_c_cases()]


@pytest.mark.parametrize("case", appendix_c_cases(), ids=CASE_IDS)
def test_appendix_c_known_answers_oracle_and_reference(case):
    bits, N, M = case["bits"], case["N"], case["M"]
    A, B = appendix_c_inputs(bits, N, M)
    engines = [Oracle(bits, N, M)] + ([Ref(bits, N, M)] if have_ref() and Ref(bits, N, M).supported() else [])
    for e in engines:
        fa, fb = e.run("fwd", A), e.run("fwd", B)
        assert [int(x) for x in fa[0, 0, :4]] == case["fwdA_head"]
        m = e.run("mul", fa, fb)
        assert (sha(fa), sha(m), sha(e.run("inv", m))) == (case["fwdA"], case["mulAB"], case["invAB"])


@pytest.mark.gpu
@pytest.mark.parametrize("case", appendix_c_cases(), ids=CASE_IDS)
def test_appendix_c_known_answers_device(case):
    bits, N, M = case["bits"], case["N"], case["M"]
    A, B = appendix_c_inputs(bits, N, M)
    c = nb.Context(bits, N, M)
    fa, fb = c.run_device("ntt_fwd", A), c.run_device("ntt_fwd", B)
    assert [int(x) for x in fa[0, 0, :4]] == case["fwdA_head"]
    m = c.run_device("mul", fa, fb)
    inv = c.run_device("ntt_inv", m)
    assert (sha(fa), sha(m), sha(inv)) == (case["fwdA"], case["mulAB"], case["invAB"])
    assert sha(c.run_device("polymul", A, B)) == case["invAB"]       # the fused product is the same four calls
    assert sha(c.host_op("fwd", A)) == case["fwdA"]                  # and so is the host-buffer entry point
    c.close()


# ---- compile-time expression programs -----------------------------------------------------------------------------------

def eval_shapes():
    out = []
    with open(os.path.join(ROOT, "nfllib_b200", "csrc", "eval_shapes.inc")) as f:
        for line in f:
            if line.startswith("NFLGPU_EVAL_SHAPE("):
                out.append([int(t, 16) for t in line[line.index("(") + 1:line.index(")")].split(",")[1:]])
    return out


def test_eval_shape_table_is_what_the_generator_writes():
    got = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_eval_shapes.py")], capture_output=True, text=True, check=True).stdout
    with open(os.path.join(ROOT, "nfllib_b200", "csrc", "eval_shapes.inc")) as f:
        assert f.read() == got
    shapes = eval_shapes()
    assert len(shapes) >= 60 and [0, 1, 2, 0x12, 0x10] in shapes and [0, 1, 2, 3, 0x13, 0x10] in shapes


def oracle_postfix(o, prog, ops):
    st = []
    for t in prog:
        if t < 8:
            st.append(ops[t])
        elif t == 0x14:
            st.append(o.run("compute_shoup", st.pop()))
        elif t == 0x13:
            yp, y, x = st.pop(), st.pop(), st.pop()
            st.append(o.run("mul_shoup", x, y, yp))
        else:
            y, x = st.pop(), st.pop()
            st.append(o.run({0x10: "add", 0x11: "sub", 0x12: "mul"}[t], x, y))
    assert len(st) == 1
    return st[0]


@pytest.mark.gpu
@pytest.mark.parametrize("bits,N,M", [(64, 1024, 4), (32, 4096, 3), (16, 512, 2)])
def test_every_compiled_expression_program_against_oracle_and_interpreter(bits, N, M, monkeypatch):
    """Each program of eval_shapes.inc: the kernel compiled for it, the interpreter (NFLGPU_EVAL_INTERPRET=1) and the oracle
    evaluating the tree one functor at a time give the same limbs.  mul_shoup consumes (y, shoup(y)) pairs, so the operand that
    plays y' is compute_shoup of the operand that plays y."""
    c, o = nb.Context(bits, N, M), Oracle(bits, N, M)
    batch = 5
    host = [random_polys(bits, N, M, batch, 700 + i) for i in range(5)]
    out = c.alloc(batch)
    for prog in eval_shapes():
        ops = list(host)
        # make every mul_shoup's third argument the Shoup word of its second when both are leaves (else skip the shape's check
        # against the oracle's mul_shoup contract and compare static vs interpreter only)
        contract_ok = True
        for i, t in enumerate(prog):
            if t == 0x13:
                if i >= 2 and prog[i - 1] < 8 and prog[i - 2] < 8:
                    ops[prog[i - 1]] = o.run("compute_shoup", ops[prog[i - 2]])
                else:
                    contract_ok = False
        nops = max(t for t in prog if t < 8) + 1
        dev = []
        for h in ops[:nops]:
            p = c.alloc(batch)
            c.upload(p, h, batch)
            dev.append(p)
        got = {}
        for mode in ("static", "interpreted"):
            if mode == "interpreted":
                monkeypatch.setenv("NFLGPU_EVAL_INTERPRET", "1")
            else:
                monkeypatch.delenv("NFLGPU_EVAL_INTERPRET", raising=False)
            c.eval(out, dev, prog, batch)
            g = np.empty_like(host[0])
            c.download(g, out, batch)
            c.sync()
            got[mode] = g
        monkeypatch.delenv("NFLGPU_EVAL_INTERPRET", raising=False)
        assert np.array_equal(got["static"], got["interpreted"]), prog
        if contract_ok:
            assert np.array_equal(got["static"], oracle_postfix(o, prog, ops)), prog
        for p in dev:
            c.free(p)
    # an operand used twice is passed as two leaves; dst may alias an operand
    a, b = host[0], host[1]
    pa, pb = c.alloc(batch), c.alloc(batch)
    c.upload(pa, a, batch)
    c.upload(pb, b, batch)
    c.eval(pa, [pa, pb], [0, 1, 0x12, 0, 0x10], batch)   # a = a*b + a
    g = np.empty_like(a)
    c.download(g, pa, batch)
    c.sync()
    assert np.array_equal(g, o.run("add", o.run("mul", a, b), a))
    for p in (pa, pb, out):
        c.free(p)
    c.close()


# ---- == / != ---------------------------------------------------------------------------------------------------------------

@pytest.mark.gpu
@pytest.mark.parametrize("bits,N,M", [(64, 1024, 4), (32, 4096, 14), (16, 512, 2), (64, 16384, 2)])
def test_device_any_equal_any_different(bits, N, M):
    """nflgpu_any_eq / nflgpu_any_neq keep expr::operator bool's ANY-coefficient semantics (ops.hpp:81-117)."""
    import torch
    c = nb.Context(bits, N, M)
    batch = 9
    a = random_polys(bits, N, M, batch, 31)
    b = a.copy()
    b[1] = random_polys(bits, N, M, 1, 32)[0]       # all different (with overwhelming probability)
    b[1, M - 1, N - 1] = a[1, M - 1, N - 1]         # ... except the very last coefficient
    b[2] = random_polys(bits, N, M, 1, 33)[0]
    b[2][b[2] == a[2]] ^= 1                          # strictly all different
    b[3, 0, 0] ^= 1                                  # all equal except the first
    b[4, M // 2, N // 2 + 1] ^= 1                    # ... except one in the middle
    pa, pb = c.alloc(batch), c.alloc(batch)
    c.upload(pa, a, batch)
    c.upload(pb, b, batch)
    flags = torch.zeros(batch, dtype=torch.uint8, device="cuda")
    c.any_eq(flags.data_ptr(), pa, pb, batch)
    c.sync()
    want_eq = np.array([(a[i] == b[i]).any() for i in range(batch)])
    assert np.array_equal(flags.cpu().numpy().astype(bool), want_eq)
    assert list(want_eq[:5]) == [True, True, False, True, True]
    c.any_neq(flags.data_ptr(), pa, pb, batch)
    c.sync()
    want_ne = np.array([(a[i] != b[i]).any() for i in range(batch)])
    assert np.array_equal(flags.cpu().numpy().astype(bool), want_ne)
    assert list(want_ne[:5]) == [False, True, True, True, True]
    c.free(pa)
    c.free(pb)
    c.close()


# ---- pooled scratch -----------------------------------------------------------------------------------------------------------

@pytest.mark.gpu
def test_scratch_pool_blocks_are_reused_and_trimmed():
    c = nb.Context(64, 1024, 4)
    p1 = c.scratch_alloc(8)
    c.scratch_free(p1)
    p2 = c.scratch_alloc(8)     # stream-ordered pool: the freed block comes back
    assert p1 == p2
    a = random_polys(64, 1024, 4, 8, 3)
    c.upload(p2, a, 8)
    c.ntt_fwd(p2, p2, 8)
    got = np.empty_like(a)
    c.download(got, p2, 8)
    c.sync()
    assert np.array_equal(got, Oracle(64, 1024, 4).run("fwd", a))
    c.scratch_free(p2)
    c.trim()
    c.close()


# ---- residues sharded over devices: the gather ----------------------------------------------------------------------------------

@pytest.mark.gpu
@pytest.mark.parametrize("bits,N,M,split", [(32, 4096, 14, (7, 7)), (64, 1024, 4, (1, 3)), (64, 8192, 6, (2, 2, 2))])
def test_gather_residues_from_slab_contexts_equals_the_full_context(bits, N, M, split):
    """Every residue group transforms its slab with a first_modulus context; nflgpu_gather_residues places the slabs in
    [batch][M][N]; the result is the full context's transform (core.hpp:594-600 is per residue)."""
    batch = 33
    a = random_polys(bits, N, M, batch, 77)
    full = nb.Context(bits, N, M)
    want = full.run_device("ntt_fwd", a)
    slabs, r0 = [], 0
    ctxs = []
    for k in split:
        cs = nb.Context(bits, N, k, first_modulus=r0)
        ctxs.append(cs)
        loc = np.ascontiguousarray(a[:, r0:r0 + k, :])
        p = cs.alloc(batch)
        cs.upload(p, loc, batch)
        cs.ntt_fwd(p, p, batch)
        cs.sync()
        slabs.append((p, r0, k))
        r0 += k
    dst = full.alloc(batch)
    full.gather_residues(dst, slabs, batch)
    got = np.empty_like(a)
    full.download(got, dst, batch)
    full.sync()
    assert np.array_equal(got, want)
    with pytest.raises(nb.NflGpuError):
        full.gather_residues(dst, [(slabs[0][0], M - 1, 2)], batch)   # residue range outside the context
    # the consumer fused with the gather: the CRT lift (gmp.hpp:183-209) reading each residue from its slab == the lift of the
    # gathered batch == the big-integer restatement
    import torch
    from oracle_lib import crt_lift
    W = full.lift_words()
    w1 = torch.zeros((batch, N, W), dtype=torch.int64, device="cuda")
    w2 = torch.zeros_like(w1)
    full.poly2mpz_slabs(w1.data_ptr(), slabs, batch)
    full.poly2mpz(w2.data_ptr(), dst, batch)
    full.sync()
    assert torch.equal(w1, w2)
    assert np.array_equal(w1[:3].cpu().numpy().view(np.uint64), crt_lift(want[:3], [int(p) for p in full.moduli]))
    with pytest.raises(nb.NflGpuError):
        full.poly2mpz_slabs(w1.data_ptr(), slabs[:-1], batch)         # a residue is missing
    for cs, (p, _, _) in zip(ctxs, slabs):
        cs.free(p)
        cs.close()
    full.free(dst)
    full.close()


def _ipc_worker(rank, conn, bits, N, M, batch, ndev):
    """One process of the two-process gather: owns residues [rank*M/2, (rank+1)*M/2), exports its slab, gathers both."""
    try:
        import nfllib_b200 as nbw
        dev = rank % ndev
        half = M // 2
        a = random_polys(bits, N, M, batch, 4040)
        cs = nbw.Context(bits, N, half, device=dev, first_modulus=rank * half)
        mine = cs.alloc(batch)
        cs.upload(mine, np.ascontiguousarray(a[:, rank * half:(rank + 1) * half, :]), batch)
        cs.ntt_fwd(mine, mine, batch)
        cs.sync()                                  # the slab is complete before its handle leaves the process
        conn.send(cs.ipc_export(mine))
        peer_handle = conn.recv()                  # the parent swaps the two handles (acts as the barrier too)
        full = nbw.Context(bits, N, M, device=dev)
        peer = full.ipc_open(peer_handle)
        dst = full.alloc(batch)
        slabs = [(mine, rank * half, half), (peer, (1 - rank) * half, half)]
        full.gather_residues(dst, slabs, batch)
        got = np.empty_like(a)
        full.download(got, dst, batch)
        full.sync()
        import torch
        W = full.lift_words()
        words = torch.zeros((batch, N, W), dtype=torch.int64, device=f"cuda:{dev}")
        lifted = torch.zeros_like(words)
        full.poly2mpz_slabs(words.data_ptr(), slabs, batch)      # the lift reads the peer's residues through the mapping
        full.poly2mpz(lifted.data_ptr(), dst, batch)
        full.sync()
        conn.send(sha(got) + (":lift-ok" if torch.equal(words, lifted) else ":lift-differs"))
        conn.recv()                                # both have finished reading: safe to unmap and free
        full.ipc_close(peer)
        conn.send("done")
    except Exception as e:  # noqa: BLE001
        conn.send(f"error: {e!r}")


@pytest.mark.gpu
def test_gather_residues_across_two_processes_over_cuda_ipc():
    """One process per residue group (both on GPU 0 when the box has a single GPU, on GPUs 0 and 1 otherwise): each exports its
    slab with nflgpu_ipc_export, maps the peer's with nflgpu_ipc_open and gathers the full RNS vector with strided copies."""
    import torch
    bits, N, M, batch = 32, 4096, 14, 64
    ndev = min(torch.cuda.device_count(), 2)
    ctx = mp.get_context("spawn")
    pipes = [ctx.Pipe() for _ in range(2)]
    procs = [ctx.Process(target=_ipc_worker, args=(r, pipes[r][1], bits, N, M, batch, ndev)) for r in range(2)]
    for p in procs:
        p.start()
    try:
        handles = [pipes[r][0].recv() if pipes[r][0].poll(180) else None for r in range(2)]
        assert all(isinstance(h, bytes) and len(h) == 64 for h in handles), handles
        pipes[0][0].send(handles[1])
        pipes[1][0].send(handles[0])
        hashes = [pipes[r][0].recv() if pipes[r][0].poll(180) else None for r in range(2)]
        for r in range(2):
            pipes[r][0].send("ok")
        a = random_polys(bits, N, M, batch, 4040)
        want = sha(Oracle(bits, N, M).run("fwd", a)) + ":lift-ok"
        assert hashes == [want, want], hashes
        assert [pipes[r][0].recv() if pipes[r][0].poll(60) else None for r in range(2)] == ["done", "done"]
    finally:
        for p in procs:
            p.join(30)
            if p.is_alive():
                p.kill()


# ---- BASELINE.json configurations at their full batch sizes ----------------------------------------------------------------------

FULL = [("C2", 64, 1024, 4, 4096), ("C3", 64, 16384, 8, 1024), ("C4", 32, 4096, 14, 8192), ("C5", 64, 8192, 6, 2048)]


@pytest.mark.gpu
@pytest.mark.parametrize("name,bits,N,M,batch", FULL, ids=[f[0] for f in FULL])
def test_baseline_configs_at_full_batch(name, bits, N, M, batch):
    """configs[1..4] of BASELINE.json with their own batch sizes (128 MiB .. 1.75 GiB per operand): the whole forward output
    against the multi-threaded unmodified reference when it travelled (else the oracle on first / last / strided
    polynomials), the round trip over the whole batch, and — C5's path — the whole fused product."""
    import torch
    c = nb.Context(bits, N, M)
    a = random_polys(bits, N, M, batch, 2024)
    pa, pf = c.alloc(batch), c.alloc(batch)
    c.upload(pa, a, batch)
    c.ntt_fwd(pf, pa, batch)
    fa = np.empty_like(a)
    c.download(fa, pf, batch)
    c.sync()
    threads = os.cpu_count() or 1
    sel = sorted(set([0, 1, batch // 2, batch - 2, batch - 1] + list(range(0, batch, max(1, batch // 61)))))
    if have_ref():
        assert np.array_equal(fa, Ref(bits, N, M).run("fwd", a, threads=threads))
    else:
        assert np.array_equal(fa[sel], Oracle(bits, N, M).run("fwd", a[sel]))
    # round trip over the whole batch, compared on the device: no coefficient may differ
    c.ntt_inv(pf, pf, batch)
    flags = torch.zeros(batch, dtype=torch.uint8, device="cuda")
    c.any_neq(flags.data_ptr(), pf, pa, batch)
    c.sync()
    assert int(flags.sum().item()) == 0
    if name == "C5":
        b = random_polys(bits, N, M, batch, 2025)
        pb = c.alloc(batch)
        c.upload(pb, b, batch)
        c.polymul(pf, pa, pb, batch)
        prod = np.empty_like(a)
        c.download(prod, pf, batch)
        c.sync()
        if have_ref():
            assert np.array_equal(prod, Ref(bits, N, M).run("polymul", a, b, threads=threads))
        else:
            assert np.array_equal(prod[sel], Oracle(bits, N, M).run("polymul", a[sel], b[sel]))
        c.free(pb)
    c.free(pa)
    c.free(pf)
    c.close()


# ---- hwt sampler: the data-dependent nonce count ---------------------------------------------------------------------------------

def test_oracle_hwt_test_knob_only_changes_the_rejection_rule():
    o = Oracle(64, 1024, 2)
    key = bytes(range(32))
    a, ca = o.hwt(6, 64, key, 5)
    b, cb = o.hwt(6, 64, key, 5, test_shrink=0)
    assert ca == cb == 6 * 16 and np.array_equal(a, b)
    c, cc = o.hwt(64, 64, key, 100, test_shrink=12)   # a rejection costs this shape a whole extra refill: 11 of 64 polys deviate
    assert cc > 64 * 16


@pytest.mark.gpu
@pytest.mark.parametrize("bits,N,M,hwt,shrink", [(32, 512, 1, 20, 7), (64, 1024, 2, 64, 12), (64, 256, 1, 16, 8), (64, 1024, 2, 64, 0)])
def test_hwt_follows_the_reference_nonce_sequence_when_a_draw_needs_an_extra_refill(bits, N, M, hwt, shrink, monkeypatch):
    """poly::set(hwt_dist) (core.hpp:355-392) refills its index buffer one more time when rejected indices push a polynomial
    past a refill boundary, which moves the start nonce of every later polynomial.  With the real rejection rule that has
    probability < 2^-44 per draw, so the test shrinks the acceptance range in the device sampler and in the oracle alike
    (NFLGPU_HWT_TEST_REJECT_SHIFT / test_shrink): 2, 11 and 36 of 64 polynomials then deviate, and the device must still
    reproduce the sequential stream — polynomials and nonce count."""
    if shrink:
        monkeypatch.setenv("NFLGPU_HWT_TEST_REJECT_SHIFT", str(shrink))
    else:
        monkeypatch.delenv("NFLGPU_HWT_TEST_REJECT_SHIFT", raising=False)
    c, o = nb.Context(bits, N, M), Oracle(bits, N, M)
    key = bytes(range(32))
    batch = 64
    want, calls = o.hwt(batch, hwt, key, 100, test_shrink=shrink)
    normal = (N - hwt + hwt - 1) // hwt + 1
    assert (calls != batch * normal) == bool(shrink)
    p = c.alloc(batch)
    used = c.hwt_count(p, batch, hwt, key, 100)
    got = np.empty_like(want)
    c.download(got, p, batch)
    c.sync()
    assert used == calls
    assert np.array_equal(got, want)
    c.free(p)
    c.close()


# ---- transforms larger than one tile: thread-block clusters ------------------------------------------------------------------------

_SPLIT_SNIPPET = """
import sys, hashlib, numpy as np
sys.path.insert(0, {root!r}); sys.path.insert(0, {tests!r})
import nfllib_b200 as nb
from oracle_lib import random_polys
c = nb.Context(64, {N}, {M})
a = random_polys(64, {N}, {M}, {batch}, 9090)
f = c.run_device("ntt_fwd", a)
print(hashlib.sha256(f.tobytes()).hexdigest(), hashlib.sha256(c.run_device("ntt_inv", a).tobytes()).hexdigest())
"""


@pytest.mark.gpu
@pytest.mark.parametrize("n,M,batch", [(15, 2, 100), (15, 5, 31)])
def test_cluster_transforms_with_more_units_than_clusters(n, M, batch):
    """64-bit N = 2^15 runs in clusters of 2 CTAs with the unit in distributed shared memory (ntt_cluster.cuh); with 200 / 155
    units the persistent clusters walk several units each.  Against the oracle on a spread of polynomials, the round
    trip over the whole batch, the fused product, and — whole batch, hash against hash — the round-1 path (global-memory pass +
    tile kernel), which NFLGPU_NO_CLUSTER=1 selects in a fresh process."""
    N = 1 << n
    c, o = nb.Context(64, N, M), Oracle(64, N, M)
    a = random_polys(64, N, M, batch, 9090)
    b = random_polys(64, N, M, batch, 9091)
    fa = c.run_device("ntt_fwd", a)
    ia = c.run_device("ntt_inv", a)
    sel = [0, 1, batch // 3, batch // 2, batch - 2, batch - 1]
    assert np.array_equal(fa[sel], o.run("fwd", a[sel]))
    assert np.array_equal(ia[sel], o.run("inv", a[sel]))
    assert np.array_equal(c.run_device("ntt_inv", fa), a)
    assert np.array_equal(c.run_device("ntt_fwd", a, inplace=True), fa)
    assert np.array_equal(c.run_device("polymul", a[:8], b[:8]), o.run("polymul", a[:8], b[:8]))
    env = dict(os.environ, NFLGPU_NO_CLUSTER="1")
    r = subprocess.run([sys.executable, "-c", _SPLIT_SNIPPET.format(root=ROOT, tests=os.path.join(ROOT, "tests"), N=N, M=M, batch=batch)],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.split() == [sha(fa), sha(ia)]
    c.close()


# ---- CUDA graphs ---------------------------------------------------------------------------------------------------------

@pytest.mark.gpu
@pytest.mark.parametrize("bits,N,M,batch", [(64, 1024, 4, 8), (32, 4096, 3, 5), (64, 2048, 2, 3)])
def test_transform_chain_is_capturable_in_a_cuda_graph(bits, N, M, batch):
    """The launches are plain stream-ordered kernel launches, so a small-batch chain -- forward(a), forward(b), product, inverse:
    the four calls of tests/nfllib_demo_main_op.cpp:31-45 -- can be captured once and replayed (launch-bound inner loops).  The
    dynamically scheduled sizes need one eager call on the capture stream first (their per-stream counter set is created then);
    replays on new operand values are bit-identical to the oracle's negacyclic product."""
    import torch
    c, o = nb.Context(bits, N, M), Oracle(bits, N, M)
    view = {32: np.int32, 64: np.int64}[bits]
    a0, b0 = random_polys(bits, N, M, batch, 4101), random_polys(bits, N, M, batch, 4102)
    a1, b1 = random_polys(bits, N, M, batch, 4103), random_polys(bits, N, M, batch, 4104)
    da, db = torch.from_numpy(a0.view(view)).cuda(), torch.from_numpy(b0.view(view)).cuda()
    fa, fb, out = torch.empty_like(da), torch.empty_like(da), torch.empty_like(da)
    s = torch.cuda.Stream()

    def chain(sh):
        c.ntt_fwd(fa.data_ptr(), da.data_ptr(), batch, sh)
        c.ntt_fwd(fb.data_ptr(), db.data_ptr(), batch, sh)
        c.mul(fa.data_ptr(), fa.data_ptr(), fb.data_ptr(), batch, sh)
        c.ntt_inv(out.data_ptr(), fa.data_ptr(), batch, sh)

    with torch.cuda.stream(s):
        chain(s.cuda_stream)  # eager: function attributes, counter sets
    s.synchronize()
    want0 = o.run("polymul", a0, b0)
    assert np.array_equal(out.cpu().numpy().view(a0.dtype), want0)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        chain(torch.cuda.current_stream().cuda_stream)
    out.zero_()
    g.replay()
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy().view(a0.dtype), want0)
    da.copy_(torch.from_numpy(a1.view(view)))
    db.copy_(torch.from_numpy(b1.view(view)))
    for _ in range(3):  # replays back to back: the dynamic walk's counters are re-armed by every launch
        g.replay()
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy().view(a0.dtype), o.run("polymul", a1, b1))
    c.close()
