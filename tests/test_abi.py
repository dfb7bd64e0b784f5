"""C-ABI surface checks that need no GPU: the library loads, exports every symbol include/nflgpu.h declares, argument
validation works, and — on a machine without a CUDA device — compute entry points fail loudly instead of falling back."""
import ctypes
import os
import re

import pytest
import torch

import nfllib_b200 as nb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    with open(os.path.join(ROOT, "include", "nflgpu.h")) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(nflgpu_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = nb.lib()
    names = declared_symbols()
    assert len(names) >= 25 and "nflgpu_ntt_fwd" in names and "nflgpu_eval" in names
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_header_is_plain_c():
    """include/nflgpu.h must compile as C (no C++/torch types in the signatures)."""
    import subprocess
    src = '#include "nflgpu.h"\nint main(void) { return (int)sizeof(nflgpu_status) * 0; }\n'
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), "-x", "c", "-", "-fsyntax-only"],
                       input=src, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_argument_validation_without_touching_a_device():
    lib = nb.lib()
    h = ctypes.c_void_p()
    assert lib.nflgpu_ctx_create(ctypes.byref(h), 24, 1024, 1, 0, 0, None, None) == -1        # limb_bits
    assert b"limb_bits" in lib.nflgpu_last_error()
    assert lib.nflgpu_ctx_create(ctypes.byref(h), 64, 1000, 1, 0, 0, None, None) == -1        # not a power of two
    assert lib.nflgpu_ctx_create(ctypes.byref(h), 16, 1024, 1, 0, 0, None, None) == -1        # > kMaxPolyDegree
    assert lib.nflgpu_ctx_create(ctypes.byref(h), 16, 512, 3, 0, 0, None, None) == -1         # > kMaxNbModuli
    assert lib.nflgpu_ctx_create(None, 64, 1024, 1, 0, 0, None, None) == -1
    assert lib.nflgpu_ntt_fwd(None, None, None, 0, None) == -1                                # null context
    assert lib.nflgpu_batch_bytes(None, 5) == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful on a machine without a GPU")
def test_no_cpu_fallback():
    with pytest.raises(nb.NflGpuError) as e:
        nb.Context(64, 1024, 4)
    assert "no CUDA device" in str(e.value) or "CUDA" in str(e.value)


def test_only_tests_bench_and_smoke_touch_the_oracle():
    """The product (nfllib_b200/, include/) must never reference oracle/ (DESIGN.md section 2)."""
    offenders = []
    for base in ("nfllib_b200", "include"):
        for d, _, files in os.walk(os.path.join(ROOT, base)):
            for fn in files:
                if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                    with open(os.path.join(d, fn), errors="ignore") as f:
                        t = f.read()
                    if re.search(r"oracle_lib|liboracle|libnflref|oracle/", t):
                        offenders.append(os.path.join(d, fn))
    assert not offenders, offenders
