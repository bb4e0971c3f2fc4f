"""CPU: gapped DP for regions with DNA wildcard columns (the case the integer kernels refuse).
  * the float restatement (oracle/mauve_oracle.c orc_nw_align_f) against golden paths minted from the reference's own NWSmall
    (tests/golden/make_golden_nw_wild.py), against the reference directly on more inputs, and against the integer restatement on
    pure ACGT input;
  * the __host__ __device__ value functions of csrc/dpwild.cu -- what every thread of nw_wild_kernel executes -- run on the CPU
    through tests/_emu.py against all of the above.
"""
import ctypes as C

import numpy as np
import pytest

import _emu
import _golden
import _oracle


def _golden_pairs():
    z = _golden.npz("nw_wild.npz")
    return [(z["a%d" % i].tobytes(), z["b%d" % i].tobytes(), z["p%d" % i].tobytes()) for i in range(int(z["n"]))]


def _emu_align(a, b):
    buf = np.zeros(len(a) + len(b) + 1, dtype=np.uint8)
    score = C.c_float(0)
    n = _emu.emu().emu_nw_wild(a, len(a), b, len(b), buf.ctypes.data, C.byref(score))
    return (buf[:n].tobytes(), float(score.value)) if n >= 0 else (None, None)


def test_float_restatement_matches_the_golden_paths():
    pairs = _golden_pairs()
    assert len(pairs) > 150
    for a, b, want in pairs:
        got, _ = _oracle.nw_align_f(a, b)
        assert got == want, (a, b)


def test_device_value_functions_match_the_golden_paths_and_the_restatement():
    for a, b, want in _golden_pairs():
        got, score = _emu_align(a, b)
        assert got == want, (a, b)
        assert score == _oracle.nw_align_f(a, b)[1]
    assert _emu_align(b"AC1", b"ACG") == (None, None)       # not a DNA letter or wildcard
    assert _emu_align(b"AC-", b"ACG") == (None, None)


def test_float_path_equals_the_integer_path_on_acgt(orc):
    rng = np.random.default_rng(3)
    for _ in range(120):
        a = bytes(rng.choice(list(b"ACGTacgt"), int(rng.integers(1, 200))).astype(np.uint8))
        b = bytes(rng.choice(list(b"ACGT"), int(rng.integers(1, 200))).astype(np.uint8))
        pi, si = orc.nw_align(a, b)
        pf, sf = _oracle.nw_align_f(a, b)
        pe, se = _emu_align(a, b)
        assert pi == pf == pe and float(si) == sf == se, (a, b)


@pytest.mark.parametrize("wild", [b"N", b"Xx", b"MRWSYKVHDBmrwsykvhdb"])
def test_float_restatement_vs_reference_more_inputs(refc, wild):
    rng = np.random.default_rng(len(wild))
    for _ in range(60):
        a = bytes(rng.choice(list(b"ACGT") * 3 + list(wild), int(rng.integers(1, 180))).astype(np.uint8))
        b = bytes(rng.choice(list(b"ACGT") * 3 + list(wild), int(rng.integers(1, 180))).astype(np.uint8))
        assert _oracle.nw_align_f(a, b)[0] == refc.nw_align(a, b)[0], (a, b)
        assert _emu_align(a, b)[0] == refc.nw_align(a, b)[0], (a, b)
