"""Static checks of the built library (CPU suite; cuobjdump only): the properties DESIGN.md §4.1 claims for the kernels are read
back from the SASS / resource usage of nfllib_b200/libnflgpu.so, so a toolchain or source change that silently loses one of them
(occupancy, TMA staging, the prefetch, the integer-multiply formulation) fails here rather than showing up as a slower bench."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "nfllib_b200", "libnflgpu.so")

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not on PATH")

FWD10 = "_ZN6nflgpu14ntt_fwd_kernelILi64ELi10ELb0EEEvNS_7NttArgsE"
INV10 = "_ZN6nflgpu14ntt_inv_kernelILi64ELi10EEEvNS_7NttArgsE"
FWD13 = "_ZN6nflgpu14ntt_fwd_kernelILi64ELi13ELb0EEEvNS_7NttArgsE"
FWD12_32 = "_ZN6nflgpu14ntt_fwd_kernelILi32ELi12ELb0EEEvNS_7NttArgsE"


def sass(fun):
    out = subprocess.run(["cuobjdump", "-sass", "-fun", fun, LIB], capture_output=True, text=True).stdout
    ops = re.findall(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)", out, flags=re.M)
    assert ops, f"{fun} not found in {LIB}"
    return ops


def resources():
    out = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    res = {}
    for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+)", out):
        res[m.group(1)] = {"reg": int(m.group(2)), "stack": int(m.group(3)), "shared": int(m.group(4))}
    return res


def test_library_is_built_for_sm_100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", LIB], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_headline_kernels_keep_their_occupancy():
    """One 896-thread CTA per SM for N = 1024 x 64-bit = 7 warps per sub-partition: at most 72 registers per thread
    (65536 / 896 = 73); spills stay marginal."""
    r = resources()
    for k in (FWD10, INV10):
        assert r[k]["reg"] <= 72, (k, r[k])
        assert r[k]["stack"] <= 128, (k, r[k])
    assert r[FWD13]["reg"] <= 128 and r[FWD13]["stack"] <= 256


def test_headline_kernels_use_tma_prefetch_and_wide_multiplies():
    for k in (FWD10, INV10):
        ops = sass(k)
        assert any(o.startswith("UBLKCP") for o in ops), "twiddle table is no longer staged by a TMA bulk copy"
        assert "SYNCS.PHASECHK.TRANS64.TRYWAIT" in ops, "mbarrier wait missing"
        assert any(o.startswith("CCTL.E.PF2") for o in ops), "L2 prefetch of the next unit missing"
        # (the two IMAD.HI of the prologue are blockIdx.x / nmoduli)
        assert sum(o.startswith("IMAD.HI") for o in ops) <= 4, "64-bit high product must stay on IMAD.WIDE (IMAD.HI is two issue slots)"
        wide = sum(o.startswith("IMAD.WIDE") for o in ops)
        # 80 butterflies per thread and unit, 6 IMAD.WIDE each (+ the N^-1 / canonicalisation multiplies)
        assert 480 <= wide <= 560, wide
        assert any(o.startswith("BAR.SYNC") for o in ops)
        assert not any(o.startswith(("HMMA", "IMMA", "UTCHMMA", "UTCIMMA", "QMMA")) for o in ops), "no tensor-core instructions on this path"
    assert any(o.startswith("LDG.E.64.CONSTANT") for o in sass(FWD10)), "forward reads its input through the non-coherent path"
    assert any(o.startswith("STG.E.128") for o in sass(FWD10)), "forward writes 16 bytes per lane"


def test_32_bit_butterflies_use_the_min_based_conditional_subtract():
    ops = sass(FWD12_32)
    assert any(o.startswith("VIADDMNMX.U32") for o in ops)
    assert sum(o.startswith("IMAD.HI") for o in ops) <= 4  # prologue division only: the butterflies' high product is an IMAD.WIDE


INV12_32 = "_ZN6nflgpu14ntt_inv_kernelILi32ELi12EEEvNS_7NttArgsE"
INV13 = "_ZN6nflgpu14ntt_inv_kernelILi64ELi13EEEvNS_7NttArgsE"
FWD14 = "_ZN6nflgpu14ntt_fwd_kernelILi64ELi14ELb0EEEvNS_7NttArgsE"
INV14 = "_ZN6nflgpu14ntt_inv_kernelILi64ELi14EEEvNS_7NttArgsE"
CLFWD15 = "_ZN6nflgpu22ntt_cluster_fwd_kernelILi64ELi15ELb0EEEvNS_11ClusterArgsE"


def test_round2_mechanisms_are_in_the_binary():
    """Round-2 claims of DESIGN §4.1 read back from the SASS: the pipelined inverse kernels copy the next unit in with cp.async (LDGSTS)
    from N = 4096 up and not for N = 1024; the N = 16384 shape moves its pass-0 window as 16-byte vectors in both directions; the
    N = 2^15 cluster kernel scatters through distributed shared memory; the C4 kernels still fit four CTAs per SM (64 registers)."""
    for k in (INV12_32, INV13, INV14):
        assert any(o.startswith("LDGSTS") for o in sass(k)), f"{k}: cp.async copy-in of the next unit missing"
    assert not any(o.startswith("LDGSTS") for o in sass(INV10))
    assert sum(o.startswith("LDG.E.128.CONSTANT") for o in sass(FWD14)) >= 16, "N = 16384 forward: pass-0 window no longer loaded as 16-byte vectors"
    assert sum(o.startswith("STG.E.128") for o in sass(INV14)) >= 16, "N = 16384 inverse: pass-0 window no longer stored as 16-byte vectors"
    assert not any(o.startswith("LDG.E.128.CONSTANT") for o in sass(FWD10)), "N = 1024 keeps 8-byte column loads (measured faster)"
    cl = sass(CLFWD15)
    assert any(o.startswith("UCGABAR_ARV") for o in cl) and any(o.startswith("UCGABAR_WAIT") for o in cl), "cluster barrier missing"
    assert sum(o.startswith("ST.E.64") for o in cl) >= 16, "pass-0 scatter into the partner CTA's tile (st.shared::cluster) missing"
    r = resources()
    for k in (FWD12_32, INV12_32):
        assert r[k]["reg"] <= 64 and r[k]["stack"] == 0, (k, r[k])


CLINV15 = "_ZN6nflgpu22ntt_cluster_inv_kernelILi64ELi15EEEvNS_11ClusterArgsE"


def test_folded_scaling_of_the_cluster_inverse_is_in_the_binary():
    """ntt_plan.h plan_fold: the N = 2^15 cluster inverse carries N^-1 in the twiddles of its first pass and multiplies x[0] once per
    thread -- 240 butterflies x 6 IMAD.WIDE + one more product -- where the other 64-bit inverse kernels multiply every sum output of
    the last stage (N = 16384: 224 butterflies + 16 of those = 240 products)."""
    wide15 = sum(o.startswith("IMAD.WIDE") for o in sass(CLINV15))
    assert 240 * 6 <= wide15 <= 240 * 6 + 24, wide15
    wide14 = sum(o.startswith("IMAD.WIDE") for o in sass(INV14))
    assert wide14 >= (224 + 16) * 6, wide14
