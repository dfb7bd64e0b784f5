"""What chunking alone costs on this box's PCIe (development aid): 2 x 128 MiB up and 2 x 128 MiB down per step, pinned memory,
(a) whole-array copies on two streams (bench.py's copy_only_ceiling), (b) chunked copies, uploads and downloads independent,
(c) chunked, each download waiting (event) for the upload of the same chunk -- the dependency structure of nflgpu_host_op without
its kernels."""
import sys, time
import torch

MiB = 1 << 20
total = 128 * MiB
hA, hD = (torch.empty(total, dtype=torch.uint8).pin_memory() for _ in range(2))
hB, hC = (torch.empty(total, dtype=torch.uint8).pin_memory() for _ in range(2))
dev = torch.empty(2 * total, dtype=torch.uint8, device="cuda")
s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, it=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(it):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / it


def whole():
    with torch.cuda.stream(s_in):
        dev[:total].copy_(hA, non_blocking=True)
        dev[total:].copy_(hD, non_blocking=True)
    with torch.cuda.stream(s_out):
        hB.copy_(dev[:total], non_blocking=True)
        hC.copy_(dev[total:], non_blocking=True)


def chunked(chunk, dependent):
    def fn():
        for src, dst, off in ((hA, hB, 0), (hD, hC, total)):
            for c in range(0, total, chunk):
                with torch.cuda.stream(s_in):
                    dev[off + c:off + c + chunk].copy_(src[c:c + chunk], non_blocking=True)
                    if dependent:
                        ev = torch.cuda.Event()
                        ev.record(s_in)
                with torch.cuda.stream(s_out):
                    if dependent:
                        s_out.wait_event(ev)
                    dst[c:c + chunk].copy_(dev[off + c:off + c + chunk], non_blocking=True)
    return fn


t = timed(whole)
print(f"whole arrays, two streams           : {t * 1e3:6.3f} ms/step  {2 * total / t / 1e9:5.1f} GB/s each way", flush=True)
for mib in (1, 2, 4, 8, 16, 32):
    for dep in (False, True):
        t = timed(chunked(mib * MiB, dep))
        print(f"chunks of {mib:2d} MiB, {'download waits for upload' if dep else 'independent              '}: {t * 1e3:6.3f} ms/step  {2 * total / t / 1e9:5.1f} GB/s each way", flush=True)
