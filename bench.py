#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native NFLlib hot path (contract: see the task statement / DESIGN.md).

Metric (BASELINE.json): forward+inverse NTT/s, N=1024, uint64, 4 RNS moduli, batched; one *transform* = one whole-
polynomial nfl::poly::ntt_pow_phi() or invntt_pow_invphi() (all 4 residues).  One *step* = `batch` forward
transforms + `batch` inverse transforms.  `value` = transforms/s with operands resident in HBM; `e2e` = the same
through the host-buffer C-ABI call (pinned host memory, H2D + kernels + D2H inside the timed region).

  python bench.py [--gpus N] [--steps K] [--warmup W]            # this framework (N>1: launch with torchrun)
  python bench.py --impl reference [--gpus N] ...                 # the reference's own CPU implementation
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "forward+inverse NTT/s (N=1024, uint64, 4 moduli, batched)"
UNIT = "transforms/s"
BITS, DEGREE, NMODULI, BATCH = 64, 1024, 4, 4096  # BASELINE.json configs[1], per GPU (weak scaling)
ROTATE = 3  # independent operand sets cycled through so that no kernel finds its input resident in L2
ALG_BYTES_PER_TRANSFORM = 2 * DEGREE * NMODULI * (BITS // 8)  # every coefficient read once, written once


def config(n_gpus):
    return {"workload": "C2: N=1024, uint64_t, 4 RNS moduli, batch=4096 polys per GPU (BASELINE.json configs[1])",
            "limb_bits": BITS, "degree": DEGREE, "nmoduli": NMODULI, "batch_per_gpu": BATCH, "global_batch": BATCH * n_gpus,
            "parallelism": f"batch-sharded x{n_gpus}, no data-path collective",
            "l2": f"rotating over {ROTATE} independent operand sets ({ROTATE * 4 * BATCH * DEGREE * NMODULI * 8 >> 20} MiB per GPU) "
                  "so every kernel reads from HBM, not from a previous kernel's L2 lines"}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu capture, if any."""
    path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f).get("fwd_kernel_dram_bytes_per_launch")
    return None


# ---- CPU side: the reference's own implementation (oracle/_ref) or, failing that, the C port -----------------

def cpu_engine():
    from oracle_lib import Oracle, Ref, have_ref
    if have_ref():
        r = Ref(BITS, DEGREE, NMODULI)
        return "reference", r, Ref.lib().nflref_build_flags().decode() + " " + Ref.arch
    return "port", Oracle(BITS, DEGREE, NMODULI), "oracle/nfl_oracle.c -O2"


def cpu_pass(kind, eng, a, work, threads):
    """one forward + one inverse over the sample `a` (in place in `work`)."""
    if kind == "reference":
        eng.run("fwd", a, threads=threads, out=work)
        eng.run("inv", work, threads=threads, out=work)
    else:
        work[...] = eng.run("inv", eng.run("fwd", a))


def cpu_sample(polys):
    from oracle_lib import aligned, random_polys
    import numpy as np
    a = aligned((polys, NMODULI, DEGREE), np.uint64)
    a[...] = random_polys(BITS, DEGREE, NMODULI, polys, 20260925)
    return a, aligned(a.shape, np.uint64)


def cpu_baseline(target_seconds=12.0):
    kind, eng, flags = cpu_engine()
    threads = usable_cpus() if kind == "reference" else 1
    polys = 512 * threads if kind == "reference" else 64
    a, work = cpu_sample(polys)
    cpu_pass(kind, eng, a, work, threads)  # warm-up (page faults, static tables)
    t0 = time.perf_counter()
    cpu_pass(kind, eng, a, work, threads)
    one = time.perf_counter() - t0
    reps = max(1, int(target_seconds / max(one, 1e-6)))
    t0 = time.perf_counter()
    for _ in range(reps):
        cpu_pass(kind, eng, a, work, threads)
    dt = time.perf_counter() - t0
    out = {"value": 2.0 * polys * reps / dt, "unit": UNIT, "cores": threads, "kind": kind,
           "sample": f"{reps} x (ntt_pow_phi + invntt_pow_invphi) over {polys} seeded polys of the same shape, {threads} host threads, "
                     f"{dt:.1f} s; build: {flags}",
           "host_cpu": host_cpu()}
    # latency of ONE polynomial (one thread), the figure to read beside e2e.single_poly_latency_us
    one, one_out = cpu_sample(1)
    lat = []
    for _ in range(300):
        t0 = time.perf_counter()
        if kind == "reference":
            eng.run("fwd", one, threads=1, out=one_out)
        else:
            eng.run("fwd", one)
        lat.append(time.perf_counter() - t0)
    lat.sort()
    out["single_poly_latency_us"] = lat[len(lat) // 2] * 1e6
    out.update(reference_ntt_perfs())
    return out


def reference_ntt_perfs():
    """The reference's own micro-benchmark, unmodified (tests/ntt_perfs.cpp built into oracle/_ref/ntt_perfs): microseconds per raw
    core::ntt of ONE residue (N=1024, uint64), one thread — the number BASELINE.json's '10x ntt_perfs' target refers to."""
    from oracle_lib import ref_dir
    exe = os.path.join(ref_dir(), "ntt_perfs")
    if not os.path.exists(exe):
        return {}
    try:
        txt = subprocess.run([exe], capture_output=True, text=True, timeout=120).stdout
    except (OSError, subprocess.TimeoutExpired):
        return {}
    res = {}
    for line in txt.splitlines():
        if "Time per NTT (lib)" in line:
            res["ntt_perfs_lib_us_per_residue_ntt"] = float(line.split(":")[1].split()[0])
        if "Time per NTT (org)" in line:
            res["ntt_perfs_org_us_per_residue_ntt"] = float(line.split(":")[1].split()[0])
    return res


def usable_cpus():
    """host threads this process may really use: affinity mask capped by the cgroup v2 cpu.max quota."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        with open("/sys/fs/cgroup/cpu.max") as f:
            quota, period = f.read().split()
        if quota != "max":
            n = max(1, min(n, int(int(quota) / int(period))))
    except (OSError, ValueError):
        pass
    return n


def host_cpu():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    kind, eng, flags = cpu_engine()
    threads = usable_cpus() if kind == "reference" else 1
    polys = 256 * threads if kind == "reference" else 32  # bounded sample per step
    a, work = cpu_sample(polys)
    for _ in range(args.warmup):
        cpu_pass(kind, eng, a, work, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_pass(kind, eng, a, work, threads)
    dt = time.perf_counter() - t0
    value = 2.0 * polys * args.steps / dt
    sample = (f"each step = (ntt_pow_phi + invntt_pow_invphi) over {polys} seeded polys (N=1024, uint64, 4 moduli), "
              f"{threads} host threads; build: {flags}; cpu: {host_cpu()}")
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "u64", "data": "synthetic", "config": config(args.gpus),
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(out))
    return 0


# ---- GPU side ------------------------------------------------------------------------------------------------------

class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(index)], stdout=self.tmp, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.tmp.flush()
        self.tmp.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.tmp.read().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); mx.append(float(parts[2]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        self.tmp.close()
        os.unlink(self.tmp.name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def bind_to_gpu_numa(local):
    """Pins this rank's host threads (and, where the box exposes more than one NUMA node, its page allocations) to the CPU set
    the driver reports for its GPU, BEFORE any pinned buffer is allocated.  Returns what was done, for the JSON line."""
    info = {"cpus": None, "numa_node": None, "mempolicy": "default"}
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1]
        allowed = sorted(set(cpus) & os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
            info["cpus"] = f"{allowed[0]}-{allowed[-1]} ({len(allowed)})"
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        with open(f"/sys/bus/pci/devices/{bus[-12:].lower()}/numa_node") as f:
            node = int(f.read())
        info["numa_node"] = node
        nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()]
        if node >= 0 and len(nodes) > 1:
            import ctypes
            mask = ctypes.c_ulong(1 << node)
            # set_mempolicy(MPOL_PREFERRED = 1, &mask, maxnode): pinned buffers allocated from now on come from the GPU's node
            if ctypes.CDLL(None, use_errno=True).syscall(238, 1, ctypes.byref(mask), 64) == 0:
                info["mempolicy"] = f"preferred node {node}"
        else:
            info["mempolicy"] = f"default ({len(nodes)} NUMA node(s) visible)"
    except Exception as e:  # noqa: BLE001 -- placement is best effort; the numbers say what it achieved
        info["error"] = repr(e)[:120]
    return info


SECONDARY = {  # BASELINE.json configs[2..4]: (bits, degree, nmoduli, batch)
    "C3": (64, 16384, 8, 1024),
    "C4": (32, 4096, 14, 8192),
    "C5": (64, 8192, 6, 2048),
}


def run_secondary(nb, torch, np, dist, world, rank, local, stream, peak):
    """BASELINE.json configs[2], [3], [4] on the same box: kernel times (CUDA events on the launch stream, 2 warm-up + 5 timed
    launches, operands 0.75 - 1.75 GiB, i.e. far larger than L2), fraction of the HBM peak on SURVEY 8d's algorithmic bytes,
    and a slice of every timed output checked against the CPU oracle.  Multi-GPU: C3 every rank its own batch (weak), C4
    residue x batch sharded with first_modulus contexts + the peer-memory gather a CRT lift would need, C5 the fixed batch
    of 2048 products split over the ranks (strong)."""
    from oracle_lib import Oracle, random_polys, golden_params
    from nfllib_b200 import sharding
    sh = stream.cuda_stream
    out = {}

    def timed(fn, iters=5, warm=2):
        for _ in range(warm):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(iters):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        if dist is not None:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def all_ok(flag):
        if dist is None:
            return bool(flag)
        t = torch.tensor([1 if flag else 0], device="cuda", dtype=torch.int32)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item())

    def dev(host):
        return torch.from_numpy(np.ascontiguousarray(host).view(np.uint8)).cuda()

    def host_of(t, dtype, shape):
        return t.cpu().numpy().view(dtype).reshape(shape)

    for name, (bits, N, M, batch) in SECONDARY.items():
        dtype = {32: np.uint32, 64: np.uint64}[bits]
        lb = bits // 8
        rec = {"limb_bits": bits, "degree": N, "nmoduli": M, "batch": batch}
        if name == "C4" and world > 1:
            shard = sharding.shard_residues(batch, M, world, rank)
            rec["sharding"] = f"{sharding.residue_groups(M, world)} residue groups x {world // sharding.residue_groups(M, world)} batch groups; " \
                              f"rank 0: polys [{shard.poly0},{shard.poly0 + shard.npolys}) x residues [{shard.res0},{shard.res0 + shard.nres})"
        elif name == "C5" and world > 1:
            shard = sharding.shard_batch(batch, M, world, rank)
            rec["sharding"] = f"batch of {batch} split over {world} ranks (strong scaling)"
        else:
            shard = sharding.Shard(0, batch, 0, M)
            if world > 1:
                rec["sharding"] = f"every rank its own batch of {batch} (weak scaling)"
        nb_, nr = shard.npolys, shard.nres
        ctx = nb.Context(bits, N, nr, device=local, first_modulus=shard.res0)
        P = golden_params(bits)["P"][shard.res0:shard.res0 + nr]
        o = Oracle(bits, N, M)
        a_full_slice = None
        a = random_polys(bits, N, nr, nb_, 7000 + 10 * rank + len(out), P=P)
        da = dev(a)
        df = torch.empty_like(da)
        units_bytes = 2 * N * nr * lb * nb_  # read once + written once per transform
        total_bytes = 2 * N * lb * (M * batch * (world if name == "C3" else 1))  # all ranks together
        ms_f = timed(lambda: ctx.ntt_fwd(df.data_ptr(), da.data_ptr(), nb_, sh))
        ctx.ntt_fwd(df.data_ptr(), da.data_ptr(), nb_, sh)
        torch.cuda.synchronize()
        sel = sorted(set([0, nb_ // 2, nb_ - 1]))
        fa = host_of(df, dtype, (nb_, nr, N))
        # the oracle works on the full residue set: embed this rank's residues at their place
        def oracle_rows(op, x_sel, other=None):
            full = np.zeros((len(sel), M, N), dtype=dtype)
            full[:, shard.res0:shard.res0 + nr, :] = x_sel
            if other is None:
                return o.run(op, full)[:, shard.res0:shard.res0 + nr, :]
            fo = np.zeros_like(full)
            fo[:, shard.res0:shard.res0 + nr, :] = other
            return o.run(op, full, fo)[:, shard.res0:shard.res0 + nr, :]
        ok = np.array_equal(fa[sel], oracle_rows("fwd", a[sel]))
        di = torch.empty_like(da)
        ms_i = timed(lambda: ctx.ntt_inv(di.data_ptr(), df.data_ptr(), nb_, sh))
        ctx.ntt_inv(di.data_ptr(), df.data_ptr(), nb_, sh)
        torch.cuda.synchronize()
        ok = ok and np.array_equal(host_of(di, dtype, (nb_, nr, N))[sel], a[sel])
        rec.update({"fwd_ms": ms_f, "inv_ms": ms_i, "fwd_gbs": total_bytes / ms_f / 1e6, "inv_gbs": total_bytes / ms_i / 1e6,
                    "frac": total_bytes / ms_f / 1e6 / (peak * world), "inv_frac": total_bytes / ms_i / 1e6 / (peak * world),
                    "transforms_per_s": 2.0 * batch * (world if name == "C3" else 1) / ((ms_f + ms_i) * 1e-3)})
        if name == "C5":
            b = random_polys(bits, N, nr, nb_, 7500 + rank, P=P)
            db = dev(b)
            ms_p = timed(lambda: ctx.polymul(di.data_ptr(), da.data_ptr(), db.data_ptr(), nb_, sh))
            ctx.polymul(di.data_ptr(), da.data_ptr(), db.data_ptr(), nb_, sh)
            torch.cuda.synchronize()
            ok = ok and np.array_equal(host_of(di, dtype, (nb_, nr, N))[sel], oracle_rows("polymul", a[sel], b[sel]))
            prod_bytes = 3 * N * M * lb * batch  # read a, read b, write c
            rec.update({"polymul_ms": ms_p, "products_per_s": batch / (ms_p * 1e-3), "polymul_gbs": prod_bytes / ms_p / 1e6,
                        "polymul_frac": prod_bytes / ms_p / 1e6 / (peak * world)})
            del db
        if name == "C4" and world > 1:
            # what a CRT lift (gmp.hpp:183-209) needs: every residue of this rank's polynomials on this device.  Each rank
            # maps the slab of its partner in the other residue group (CUDA IPC) and pulls it over NVLink with one strided
            # copy-engine transfer, its own slab with another (nflgpu_gather_residues).
            mine = ctx.alloc(nb_)  # an allocation of its own, so that its CUDA IPC handle names exactly this slab
            ctx.ntt_fwd(mine, da.data_ptr(), nb_, sh)
            ctx.sync(sh)
            full_ctx = nb.Context(bits, N, M, device=local)
            dst, slabs, close_peers = sharding.gather_residues_peer(full_ctx, ctx, mine, shard, world, rank, sh)
            mapped = slabs[1:]
            handles = [None] * world
            dist.all_gather_object(handles, (None, shard.poly0, shard.npolys, shard.res0, shard.nres))
            peers = [(r, handles[r]) for r in sharding.residue_partners([sharding.Shard(*h[1:]) for h in handles], rank)]
            dist.barrier()
            ms_g = timed(lambda: full_ctx.gather_residues(dst, slabs, nb_, sh), iters=10, warm=3)
            ms_peer = timed(lambda: full_ctx.gather_residues(dst, mapped, nb_, sh), iters=10, warm=2)  # the NVLink part alone
            full_ctx.gather_residues(dst, slabs, nb_, sh)
            full_ctx.sync(sh)
            got = np.empty((nb_, M, N), dtype=dtype)
            full_ctx.download(got, dst, nb_)
            full_ctx.sync()
            a_sel_full = np.zeros((len(sel), M, N), dtype=dtype)
            # inputs of the other residue groups are re-derived from their seeds: same generator, their rank ids
            for r, h in [(rank, handles[rank])] + peers:
                Pr = golden_params(bits)["P"][h[3]:h[3] + h[4]]
                ar = random_polys(bits, N, h[4], h[2], 7000 + 10 * r + len(out), P=Pr)
                a_sel_full[:, h[3]:h[3] + h[4], :] = ar[sel]
            ok = ok and np.array_equal(got[sel], o.run("fwd", a_sel_full))
            pulled = sum(nb_ * k * N * lb for _, _, k in mapped)
            rec["gather"] = {"ms": ms_g, "peer_only_ms": ms_peer, "peer_bytes_per_rank": pulled, "peer_gbs_per_rank": pulled / ms_peer / 1e6,
                             "aggregate_peer_gbs": pulled * world / ms_peer / 1e6, "nvlink_peak_gbs_per_direction": 900.0,
                             "frac_of_nvlink": pulled / ms_peer / 1e6 / 900.0,
                             "full_vector_gbs_per_rank": nb_ * M * N * lb / ms_g / 1e6,
                             "how": "nflgpu_ipc_export/open + nflgpu_gather_residues: one strided cudaMemcpy2DAsync per slab, destination "
                                    "[batch][14][4096] written in place, the local slab on a side stream; ms = whole gather, peer_only_ms = the peer slab alone "
                                    "(what peer_gbs / frac_of_nvlink use); all ranks at once; max over ranks"}
            dist.barrier()
            close_peers()
            dist.barrier()
            full_ctx.free(dst)
            ctx.free(mine)
            full_ctx.close()
        rec["checked_vs_oracle"] = all_ok(ok)
        out[name] = rec
        del da, df, di
        ctx.close()
        torch.cuda.empty_cache()
    return out


def run_b200(args):
    import numpy as np
    import torch
    import nfllib_b200 as nb
    from oracle_lib import Oracle, random_polys

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and world == 1:
        # convenience: relaunch ourselves the way the driver does
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", str(29500 + os.getpid() % 1000), os.path.abspath(__file__), "--gpus", str(args.gpus), "--steps",
               str(args.steps), "--warmup", str(args.warmup)] + (["--no-cpu"] if args.no_cpu else []) + (["--headline-only"] if args.headline_only else [])
        return subprocess.call(cmd)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this framework has no CPU fallback (use --impl reference for the CPU baseline)")
    numa = bind_to_gpu_numa(local)  # before the first pinned allocation
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    ctx = nb.Context(BITS, DEGREE, NMODULI, device=local)
    stream = torch.cuda.current_stream()
    sh = stream.cuda_stream
    shape = (BATCH, NMODULI, DEGREE)

    def dev(host):
        return torch.from_numpy(np.ascontiguousarray(host).view(np.int64)).cuda()

    # operands: ROTATE coefficient-domain sets A, ROTATE NTT-domain sets D (forward images of other random polys)
    A, Bf, D, C = [], [], [], []
    for r in range(ROTATE):
        A.append(dev(random_polys(BITS, DEGREE, NMODULI, BATCH, 1000 * rank + 10 + r)))
        Bf.append(torch.empty(shape, dtype=torch.int64, device="cuda"))
        d = dev(random_polys(BITS, DEGREE, NMODULI, BATCH, 1000 * rank + 20 + r))
        ctx.ntt_fwd(d.data_ptr(), d.data_ptr(), BATCH, sh)
        D.append(d)
        C.append(torch.empty(shape, dtype=torch.int64, device="cuda"))
    torch.cuda.synchronize()

    def step(i, evs=None):
        r = i % ROTATE
        if evs:
            evs[0].record(stream)
        ctx.ntt_fwd(Bf[r].data_ptr(), A[r].data_ptr(), BATCH, sh)
        if evs:
            evs[1].record(stream)
        ctx.ntt_inv(C[r].data_ptr(), D[r].data_ptr(), BATCH, sh)
        if evs:
            evs[2].record(stream)

    # correctness of what is being timed: a slice against the CPU oracle (checker only, outside the timed region)
    step(0)
    torch.cuda.synchronize()
    o = Oracle(BITS, DEGREE, NMODULI)
    a0 = A[0][:2].cpu().numpy().view(np.uint64)
    checked = bool(np.array_equal(Bf[0][:2].cpu().numpy().view(np.uint64), o.run("fwd", a0))) and \
        bool(np.array_equal(C[0][:2].cpu().numpy().view(np.uint64), o.run("inv", D[0][:2].cpu().numpy().view(np.uint64))))
    if not checked:
        raise SystemExit("bench.py: GPU results differ from the oracle; refusing to report a number")

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for i in range(args.warmup):
        step(i)
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = ctx.launch_count
    t_begin.record(stream)
    for i in range(args.steps):
        step(i, evs[i])
    t_end.record(stream)
    barrier()
    launches = ctx.launch_count - launches0
    ms = max_over_ranks(t_begin.elapsed_time(t_end))
    fwd_ms = sum(e[0].elapsed_time(e[1]) for e in evs) / args.steps
    inv_ms = sum(e[1].elapsed_time(e[2]) for e in evs) / args.steps
    value = 2.0 * BATCH * world * args.steps / (ms * 1e-3)

    # ---- end to end: host buffers, H2D + kernels + D2H inside the library call (nflgpu_host_op) ----
    hA = torch.from_numpy(random_polys(BITS, DEGREE, NMODULI, BATCH, 1000 * rank + 30).view(np.int64)).pin_memory()
    hD = D[0].cpu().pin_memory()
    hB = torch.empty(shape, dtype=torch.int64).pin_memory()
    hC = torch.empty(shape, dtype=torch.int64).pin_memory()
    nA, nD, nB, nC = (t.numpy().view(np.uint64) for t in (hA, hD, hB, hC))

    def timed_host(fn, steps, tail=None):
        for _ in range(max(1, min(args.warmup, 3))):
            fn()
        if tail:
            tail()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        if tail:
            tail()  # (inside the timed region: every result is in host memory when the clock stops)
        barrier()
        return max_over_ranks(time.perf_counter() - t0)

    def e2e_step():  # one step = both batches through the asynchronous call; like the device-resident figure, the K steps are
        ctx.host_op("fwd", nA, out=nB, wait=False)  # queued back to back and waited for once, inside the timed region
        ctx.host_op("inv", nD, out=nC, wait=False)

    def e2e_synced_step():  # the same with one wait per step
        e2e_step()
        ctx.host_sync()

    def e2e_blocking_step():  # the same step as two blocking calls (each waits for its own last download before the next upload starts)
        ctx.host_op("fwd", nA, out=nB)
        ctx.host_op("inv", nD, out=nC)

    e2e_s = timed_host(e2e_step, args.steps, tail=ctx.host_sync)
    e2e_ok = bool(np.array_equal(nB[:2], o.run("fwd", nA[:2]))) and bool(np.array_equal(nB[-2:], o.run("fwd", nA[-2:]))) and \
        bool(np.array_equal(nC[-2:], o.run("inv", nD[-2:])))
    nB[:] = 0
    nC[:] = 0
    e2e_blk_s = timed_host(e2e_blocking_step, args.steps)
    e2e_sync_s = timed_host(e2e_synced_step, args.steps)
    e2e_blk_ok = bool(np.array_equal(nB[-2:], o.run("fwd", nA[-2:]))) and bool(np.array_equal(nC[:2], o.run("inv", nD[:2])))
    clocks = sampler.stop() if sampler else None  # sampled across both timed regions (device-resident + end-to-end)
    e2e_value = 2.0 * BATCH * world * args.steps / e2e_s
    poly_bytes = DEGREE * NMODULI * 8

    # the same calls on PAGEABLE host arrays (what an array of nfl::poly from posix_memalign is): staged through the
    # library's pinned buffers with one extra host memcpy each way
    pA, pD = np.array(nA), np.array(nD)
    pB, pC = np.empty_like(pA), np.empty_like(pD)

    def e2e_pageable_step():
        ctx.host_op("fwd", pA, out=pB)
        ctx.host_op("inv", pD, out=pC)

    pg_steps = max(3, args.steps // 4)
    e2e_pg_s = timed_host(e2e_pageable_step, pg_steps)
    e2e_pg_ok = bool(np.array_equal(pB[:2], o.run("fwd", pA[:2])))

    # ... and after page-locking those same arrays once (nflgpu_host_register = cudaHostRegister): direct DMA, no staging
    t0 = time.perf_counter()
    for arr in (pA, pD, pB, pC):
        ctx.host_register(arr)
    register_ms = (time.perf_counter() - t0) * 1e3
    e2e_reg_s = timed_host(e2e_pageable_step, pg_steps)
    e2e_reg_ok = bool(np.array_equal(pB[:2], o.run("fwd", pA[:2])))
    for arr in (pA, pD, pB, pC):
        ctx.host_unregister(arr)

    # copy-only ceiling of this box, same process, same bytes per step as e2e (2 x 128 MiB each way), both directions at once
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()

    def copy_step():
        with torch.cuda.stream(s_in):
            A[0].copy_(hA, non_blocking=True)
            D[1].copy_(hD, non_blocking=True)
        with torch.cuda.stream(s_out):
            hB.copy_(Bf[0], non_blocking=True)
            hC.copy_(C[0], non_blocking=True)

    def copy_tail():
        s_in.synchronize()
        s_out.synchronize()

    copy_s = timed_host(copy_step, args.steps, tail=copy_tail)
    ceiling = 2.0 * BATCH * world * args.steps / copy_s

    # latency of ONE polynomial through the host-buffer call (what poly::ntt_pow_phi() on a host poly costs)
    one_in, one_out = np.array(nA[:1]), np.empty_like(nA[:1])
    for _ in range(20):
        ctx.host_op("fwd", one_in, out=one_out)
    lat = []
    for _ in range(200):
        t0 = time.perf_counter()
        ctx.host_op("fwd", one_in, out=one_out)
        lat.append(time.perf_counter() - t0)
    lat.sort()
    single_ok = bool(np.array_equal(one_out, o.run("fwd", one_in)))

    del hA, hB, hC, hD, A, Bf, C, D
    torch.cuda.empty_cache()
    peak, peak_src = peaks()
    secondary = None if args.headline_only else run_secondary(nb, torch, np, dist, world, rank, local, stream, peak)

    if rank == 0:
        achieved = ALG_BYTES_PER_TRANSFORM * BATCH / (fwd_ms * 1e-3) / 1e9
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
               "data": "synthetic", "config": config(world),
               "roofline": {"bound": "hbm", "kernel": "ntt_fwd_kernel<64,10,false> (forward, one launch per batch)", "achieved": achieved,
                            "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic(),
                            "traffic_source": "committed ncu --set full capture (profiles/roofline_traffic.json), not measured by this run",
                            "peak_source": peak_src,
                            "algorithmic_bytes_per_launch": ALG_BYTES_PER_TRANSFORM * BATCH, "fwd_ms_per_launch": fwd_ms,
                            "inv_ms_per_launch": inv_ms,
                            "inv_achieved": ALG_BYTES_PER_TRANSFORM * BATCH / (inv_ms * 1e-3) / 1e9},
               "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 2 * BATCH * poly_bytes, "d2h_bytes_per_step": 2 * BATCH * poly_bytes,
                       "api": "per step nflgpu_host_op_async(fwd) + nflgpu_host_op_async(inv) on pinned host buffers (16 MiB chunks through a ring "
                              "of device buffers: upload, kernel and download of every chunk inside the call); the K steps are queued back to "
                              "back and nflgpu_host_sync is called once, inside the timed region, like the device-resident figure",
                       "checked_vs_oracle": e2e_ok,
                       "one_wait_per_step": {"value": 2.0 * BATCH * world * args.steps / e2e_sync_s, "unit": UNIT,
                                             "what": "the same step followed by nflgpu_host_sync every step"},
                       "blocking_calls": {"value": 2.0 * BATCH * world * args.steps / e2e_blk_s, "unit": UNIT, "checked_vs_oracle": e2e_blk_ok,
                                          "what": "the same step as nflgpu_host_op(fwd); nflgpu_host_op(inv): each call waits for its own last download"},
                       "pageable": {"value": 2.0 * BATCH * world * pg_steps / e2e_pg_s, "unit": UNIT, "checked_vs_oracle": e2e_pg_ok,
                                    "what": "the same two calls on pageable numpy arrays (the layout of posix_memalign'ed nfl::poly[]): "
                                            "staged through the library's pinned buffers by a few host threads"},
                       "pageable_registered": {"value": 2.0 * BATCH * world * pg_steps / e2e_reg_s, "unit": UNIT, "checked_vs_oracle": e2e_reg_ok,
                                               "register_ms_once": register_ms,
                                               "what": "the same arrays after one nflgpu_host_register each (4 x 128 MiB page-locked in "
                                                       "register_ms_once, outside the timed region): direct DMA"},
                       "copy_only_ceiling": {"value": ceiling, "unit": UNIT, "frac_reached": e2e_value / ceiling,
                                             "what": "cudaMemcpyAsync of the same bytes per step, H2D and D2H on two streams at once, "
                                                     "no kernel, queued for all K steps and waited for once like e2e, all ranks together: what the box's PCIe / host memory allows"},
                       "single_poly_latency_us": {"median": lat[len(lat) // 2] * 1e6, "p10": lat[len(lat) // 10] * 1e6,
                                                  "p90": lat[len(lat) * 9 // 10] * 1e6, "checked_vs_oracle": single_ok,
                                                  "what": "nflgpu_host_op(fwd, batch=1) on a pageable 32 KiB poly, rank 0"},
                       "numa": numa},
               "gpu_launches": int(launches), "clocks": clocks, "checked_vs_oracle": checked}
        if secondary is not None:
            out["configs"] = secondary
            if "gather" in secondary.get("C4", {}):
                out["collective"] = secondary["C4"]["gather"]
        if world == 1 and not args.no_cpu:
            out["cpu_baseline"] = cpu_baseline()
        print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--headline-only", action="store_true", help="skip the configs[2..4] table (profiling runs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 1)
    sys.exit(run_reference(args) if args.impl == "reference" else run_b200(args))


if __name__ == "__main__":
    main()
