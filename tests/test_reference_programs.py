"""The REFERENCE's own test programs (tests/nfl_add.cpp, nfl_sub, nfl_mul, nfl_eq, nfl_neq, nfl_stream, poly_p, poly_set,
poly_serialize_manually for each configuration of its tests/CMakeLists.txt:1-7, and tests/ntt_perfs.cpp) compiled UNCHANGED
against include/nfl_b200.hpp by tests/cpp/Makefile and run on the GPU: what a user switching from <nfl.hpp> would see.
The binaries are built where /root/reference exists (this container) and travel to the GPU box; the sources never enter
the repo.  Their own pass criterion is exit status 0 (several of them rely on the reference's any-equal `operator==`,
ops.hpp:81-95 — the strong, oracle-checked comparisons are in test_cpp_dropin.py / test_gpu_parity.py)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFDIR = os.path.join(ROOT, "tests", "cpp", "_ref")
PROGRAMS = ["nfl_add", "nfl_sub", "nfl_mul", "nfl_eq", "nfl_neq", "nfl_stream", "poly_p", "poly_set", "poly_serialize_manually"]
CONFIGS = ["8_60_uint32_t", "128_14_uint16_t", "1024_60_uint32_t", "8192_124_uint64_t", "32768_124_uint64_t"]
ALL = [p + c for p in PROGRAMS for c in CONFIGS] + ["ntt_perfs"]

# The reference's demo programs (tests/nfllib_demo_main_{op,func}.cpp: every operator, every sampler incl. the Gaussian, the
# SIMD-vs-serial mulmod_shoup check through ops::make_op :61-87, and the LWE encrypt / decrypt self-check :260-331, which
# abort the program when they fail) and its multiple-definition check (multi0.cpp + multi1.cpp) compile and link UNCHANGED
# against the drop-in header and run on the device (first run: profiles/r02a_reference_demos.txt).
DEMOS = [d + c for d in ("nfllib_demo_main_op", "nfllib_demo_main_func") for c in ("1024_60_uint32_t", "8192_124_uint64_t")] + ["ntt_multi"]

needs_build = pytest.mark.skipif(not os.path.isdir(REFDIR), reason="tests/cpp/_ref not built (needs /root/reference at build time)")


@needs_build
def test_reference_programs_are_built_against_the_c_abi_only():
    missing = [b for b in ALL if not os.path.exists(os.path.join(REFDIR, b))]
    assert not missing, f"run __graft_entry__.build(): {missing}"
    out = subprocess.run(["ldd", os.path.join(REFDIR, "ntt_perfs")], capture_output=True, text=True).stdout
    assert "libnflgpu.so" in out and "torch" not in out and "gmp" not in out


@needs_build
def test_reference_demo_programs_compile_and_link_unchanged():
    missing = [b for b in DEMOS if not os.path.exists(os.path.join(REFDIR, b))]
    assert not missing, f"run __graft_entry__.build(): {missing}"
    for b in DEMOS[:-1]:  # (ntt_multi only includes the header twice: it references no symbol at all)
        out = subprocess.run(["ldd", os.path.join(REFDIR, b)], capture_output=True, text=True).stdout
        assert "libnflgpu.so" in out and "gmp" not in out and "mpfr" not in out, b  # MPFR is opened by libnflgpu at run time only


@needs_build
def test_reference_programs_fail_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = subprocess.run([os.path.join(REFDIR, "nfl_add1024_60_uint32_t")], capture_output=True, text=True, timeout=120)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr  # no silent host path


@pytest.mark.gpu
@needs_build
@pytest.mark.parametrize("binary", ALL[:-1])
def test_reference_program_passes(binary):
    r = subprocess.run([os.path.join(REFDIR, binary)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout + r.stderr)[-2000:]


@pytest.mark.gpu
@needs_build
@pytest.mark.parametrize("binary", DEMOS)
def test_reference_demo_program_passes(binary):
    """Exit status 0 means the demo's own checks held: tests/nfllib_demo_main_op.cpp:61-87 compares mulmod_shoup through the
    explicit functor with the operator path coefficient by coefficient, :260-331 encrypts and decrypts an LWE sample and
    requires the error to vanish; both print a message and return nonzero otherwise."""
    r = subprocess.run([os.path.join(REFDIR, binary)], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, (r.stdout + r.stderr)[-2000:]
    if binary != "ntt_multi":
        assert "Time per LWE-like symmetric decryption" in r.stdout or "Time per polynomial multiplication" in r.stdout, r.stdout[-500:]


@pytest.mark.gpu
@needs_build
def test_reference_ntt_perfs_runs_on_the_device():
    """tests/ntt_perfs.cpp: 50 000 calls of poly::core::ntt through the friend proxy, one host residue per call (so this
    times PCIe round trips, not the kernel — the throughput figure is bench.py's)."""
    r = subprocess.run([os.path.join(REFDIR, "ntt_perfs")], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, (r.stdout + r.stderr)[-2000:]
    m = re.search(r"Time per NTT \(lib\): ([0-9.e+-]+) us", r.stdout)
    assert m and float(m.group(1)) > 0, r.stdout
    print(r.stdout)
