"""Host-side parameter derivation (nfllib_b200/csrc/params.cpp) against NFLlib's tables
(include/nfl/params.hpp:12-119): committed fixture everywhere, all 2 + 291 + 1000 entries when the reference
has been compiled (oracle/_ref).  No GPU needed: nflgpu_params* are pure host functions."""
import numpy as np
import pytest

from oracle_lib import Ref, have_ref, golden_params
import nfllib_b200 as nb


@pytest.mark.parametrize("bits", [16, 32, 64])
def test_params_match_golden(bits):
    g = golden_params(bits)
    n = len(g["P"])
    mine = nb.params(bits, 0, n)
    for key in ("P", "Pn", "roots", "invkmax"):
        assert [int(v) for v in mine[key]] == g[key], key
    lim = nb.params_limits(bits)
    assert lim["kmax"] == g["kmax"] and lim["maxmoduli"] == g["maxmoduli"]
    assert lim["modulus_bits"] == bits - 2


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref/libnflref.so not built (needs /root/reference)")
@pytest.mark.parametrize("bits", [16, 32, 64])
def test_params_match_reference_full_tables(bits):
    lim = nb.params_limits(bits)
    live = Ref.params(bits, lim["maxmoduli"])
    assert len(live["P"]) == lim["maxmoduli"]
    mine = nb.params(bits, 0, lim["maxmoduli"])
    for key in ("P", "Pn", "roots", "invkmax"):
        assert [int(v) for v in mine[key]] == live[key], key


def test_params_offset_window():
    a = nb.params(32, 0, 20)
    b = nb.params(32, 5, 10)
    assert np.array_equal(a["P"][5:15], b["P"]) and np.array_equal(a["roots"][5:15], b["roots"])


def test_params_out_of_range():
    with pytest.raises(nb.NflGpuError):
        nb.params(16, 0, 3)
    with pytest.raises(nb.NflGpuError):
        nb.params_limits(8)
