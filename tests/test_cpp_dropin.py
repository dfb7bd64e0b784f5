"""The C++11 drop-in header (include/nfl_b200.hpp) exercised by tests/cpp/test_dropin.cpp — a program written
against the nfl::poly surface the way the reference's own tests are — and its dumped results compared
bit-for-bit with the CPU oracle."""
import os
import subprocess

import numpy as np
import pytest

from oracle_lib import Oracle, random_polys, DTYPES

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "test_dropin")
CFG = {"u64": (64, 1024, 4), "u32": (32, 4096, 3), "u16": (16, 512, 2)}


def test_dropin_binary_is_built_and_links_only_the_c_abi():
    assert os.path.exists(BIN), "run __graft_entry__.build()"
    out = subprocess.run(["ldd", BIN], capture_output=True, text=True).stdout
    assert "libnflgpu.so" in out and "torch" not in out


@pytest.mark.gpu
@pytest.mark.parametrize("limb", ["u64", "u32", "u16"])
def test_dropin_program_matches_oracle(limb, tmp_path):
    bits, N, M = CFG[limb]
    count = 5
    a = random_polys(bits, N, M, count, 61)
    b = random_polys(bits, N, M, count, 62)
    fa, fb, fo = (str(tmp_path / n) for n in ("a.bin", "b.bin", "out.bin"))
    a.tofile(fa)
    b.tofile(fb)
    r = subprocess.run([BIN, limb, fa, fb, str(count), fo], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    got = np.fromfile(fo, dtype=DTYPES[bits])
    o = Oracle(bits, N, M)
    fwd_a, bs = o.run("fwd", a), o.run("compute_shoup", b)
    single = [fwd_a, o.run("inv", a), o.run("add", a, b), o.run("sub", a, b), o.run("mul", a, b), bs, o.run("mul_shoup", a, b, bs),
              o.run("polymul", a, b)]
    exp_single = np.stack(single, axis=1).reshape(-1)  # per poly: the 8 results in order
    batch = [fwd_a, a, o.run("mul", a, b), o.run("mul_shoup", a, b, bs), o.run("muladd", a, b, a), o.run("polymul", a, b)]
    exp = np.concatenate([exp_single] + [x.reshape(-1) for x in batch])
    assert got.size == exp.size
    assert np.array_equal(got, exp)


STRESS = os.path.join(ROOT, "tests", "cpp", "sched_stress")


@pytest.mark.gpu
def test_scheduler_stress_program_without_torch():
    """tests/cpp/sched_stress.cpp: hundreds of forward / inverse launches of the dynamically scheduled and the split sizes on two
    non-blocking streams, through the C ABI only (the program compute-sanitizer racecheck runs: profiles/r02c_sanitizers.txt)."""
    assert os.path.exists(STRESS), "run __graft_entry__.build()"
    r = subprocess.run([STRESS, "200"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "MISMATCH" not in r.stdout, r.stdout + r.stderr
