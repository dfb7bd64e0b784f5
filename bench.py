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
        return "reference", r, Ref.lib().nflref_build_flags().decode()
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
    out.update(reference_ntt_perfs())
    return out


def reference_ntt_perfs():
    """The reference's own micro-benchmark, unmodified (tests/ntt_perfs.cpp built into oracle/_ref/ntt_perfs): microseconds per raw
    core::ntt of ONE residue (N=1024, uint64), one thread — the number BASELINE.json's '10x ntt_perfs' target refers to."""
    exe = os.path.join(ROOT, "oracle", "_ref", "ntt_perfs")
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
               str(args.steps), "--warmup", str(args.warmup)]
        return subprocess.call(cmd)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this framework has no CPU fallback (use --impl reference for the CPU baseline)")
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
    ms = t_begin.elapsed_time(t_end)
    fwd_ms = sum(e[0].elapsed_time(e[1]) for e in evs) / args.steps
    inv_ms = sum(e[1].elapsed_time(e[2]) for e in evs) / args.steps
    if dist is not None:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = 2.0 * BATCH * world * args.steps / (ms * 1e-3)

    # ---- end to end: host buffers in pinned memory, H2D + kernels + D2H inside the library call ----
    hA = torch.from_numpy(random_polys(BITS, DEGREE, NMODULI, BATCH, 1000 * rank + 30).view(np.int64)).pin_memory()
    hD = D[0].cpu().pin_memory()
    hB = torch.empty(shape, dtype=torch.int64).pin_memory()
    hC = torch.empty(shape, dtype=torch.int64).pin_memory()
    nA, nD, nB, nC = (t.numpy().view(np.uint64) for t in (hA, hD, hB, hC))

    def e2e_step():
        ctx.host_op("fwd", nA, out=nB)
        ctx.host_op("inv", nD, out=nC)

    for _ in range(max(1, min(args.warmup, 3))):
        e2e_step()
    e2e_ok = bool(np.array_equal(nB[:2], o.run("fwd", nA[:2])))
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop() if sampler else None  # sampled across both timed regions (device-resident + end-to-end)
    if dist is not None:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = 2.0 * BATCH * world * args.steps / e2e_s
    poly_bytes = DEGREE * NMODULI * 8

    if rank == 0:
        peak, peak_src = peaks()
        achieved = ALG_BYTES_PER_TRANSFORM * BATCH / (fwd_ms * 1e-3) / 1e9
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
               "data": "synthetic", "config": config(world),
               "roofline": {"bound": "hbm", "kernel": "ntt_fwd_kernel<64,10,false> (forward, one launch per batch)", "achieved": achieved,
                            "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic(), "peak_source": peak_src,
                            "algorithmic_bytes_per_launch": ALG_BYTES_PER_TRANSFORM * BATCH, "fwd_ms_per_launch": fwd_ms,
                            "inv_ms_per_launch": inv_ms,
                            "inv_achieved": ALG_BYTES_PER_TRANSFORM * BATCH / (inv_ms * 1e-3) / 1e9},
               "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 2 * BATCH * poly_bytes, "d2h_bytes_per_step": 2 * BATCH * poly_bytes,
                       "api": "nflgpu_host_op(fwd) + nflgpu_host_op(inv) on pinned host buffers", "checked_vs_oracle": e2e_ok},
               "gpu_launches": int(launches), "clocks": clocks, "checked_vs_oracle": checked}
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
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 1)
    sys.exit(run_reference(args) if args.impl == "reference" else run_b200(args))


if __name__ == "__main__":
    main()
