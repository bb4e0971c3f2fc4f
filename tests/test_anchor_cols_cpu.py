"""Anchor columns of alignment windows (muscle::FindAnchorColsPP, MU/anchoredpp.cpp:354-409) on the CPU:

  * the restatement (oracle/mauve_oracle.c: orc_anchor_cols) against the goldens the REFERENCE's own functions produced
    (tests/golden/anchor_cols.npz, minted by tests/golden/make_golden_anchor_cols.py): anchor columns, per-column scores and smoothed
    scores, float for float -- and against the reference itself on fresh random windows where oracle/_ref is present;
  * the value path of the CUDA source (csrc/anchorcols.cu: ac_window compiled for the host by tests/_emu.py, one thread standing in
    for the CTA) against the same goldens.  The parallel form proper is checked on the GPU (tests/test_zzzz_next_rows_gpu.py).
"""
import ctypes as C
import os

import numpy as np
import pytest

import _emu
import _oracle
from mauve_py_b200 import synth

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "anchor_cols.npz")


def _windows():
    z = np.load(GOLDEN)
    for k in range(int(z["n_windows"])):
        yield k, z["w%d_rows" % k], int(z["w%d_n1" % k]), z["w%d_weights" % k], z["w%d_cols" % k], z["w%d_score" % k], z["w%d_smooth" % k]


def _same_floats(a, b):
    return np.array_equal(np.asarray(a, dtype=np.float32).view(np.uint32), np.asarray(b, dtype=np.float32).view(np.uint32))


def test_default_parameters_are_the_references_settings():
    z = np.load(GOLDEN)
    s, letters = z["settings"], z["letters"]
    p = _oracle.anchor_default_params()
    assert _same_floats(np.array(p.subst[:]), s[:16])
    assert (p.gap_open, p.gap_extend, p.term_gap, p.smooth_ceil, p.min_best_col, p.min_smooth) == tuple(float(x) for x in s[16:22])
    assert (p.smooth_window, p.anchor_spacing) == (int(s[22]), int(s[23])) and int(s[24]) == 4
    mine = np.array(p.letter_of_char[:], dtype=np.uint8)
    # residues and gaps agree exactly; everything else only has to lie outside the alphabet on both sides
    assert np.array_equal(mine < 4, letters < 4) and np.array_equal(mine[mine < 4], letters[letters < 4])
    assert np.array_equal(mine == 255, letters == 255)


def test_oracle_equals_golden():
    n = 0
    for k, rows, n1, w, cols, score, smooth in _windows():
        c2, s2, m2, _, _ = _oracle.anchor_cols(rows, n1, weights=w)
        assert np.array_equal(c2, cols), k
        assert _same_floats(s2, score) and _same_floats(m2, smooth), k
        n += cols.size
    assert n > 300


def _emu_cols(rows, n1, w, p=None):
    rows = np.ascontiguousarray(rows, dtype=np.uint8)
    nr, ncol = rows.shape
    cols = np.zeros(ncol + 1, dtype=np.uint32)
    score = np.zeros(ncol + 1, dtype=np.float32)
    smooth = np.zeros(ncol + 1, dtype=np.float32)
    w = np.ascontiguousarray(w, dtype=np.float32)
    p = p if p is not None else _oracle.anchor_default_params()
    n = _emu.emu().emu_anchor_cols(rows.ctypes.data, n1, nr - n1, ncol, w.ctypes.data, C.addressof(p), cols.ctypes.data, score.ctypes.data, smooth.ctypes.data)
    assert n >= 0
    return cols[:n], score[:ncol], smooth[:ncol]


def _seg_counters():
    c = np.zeros(2, dtype=np.uint64)
    _emu.emu().emu_anchor_counters(c.ctypes.data)
    return int(c[0]), int(c[1])


def test_kernel_value_path_equals_golden():
    f0, s0 = _seg_counters()
    for k, rows, n1, w, cols, score, smooth in _windows():
        c2, s2, m2 = _emu_cols(rows, n1, w)
        assert np.array_equal(c2, cols), k
        assert _same_floats(s2, score) and _same_floats(m2, smooth), k
    # both forms of the smoothing chain were at work: segments finished in exact arithmetic, and segments run as the float chain
    f1, s1 = _seg_counters()
    assert f1 - f0 > 100 and s1 - s0 > 300, (f1 - f0, s1 - s0)


def test_smoothing_segments_that_must_take_the_chain():
    """totals that lose low bits on the way (a window sum crossing a power of two with a residue on board) and operands outside the
    fixed-point range: the exact path has to refuse them, and the answer stays the oracle's"""
    rng = np.random.default_rng(17)
    p = _oracle.anchor_default_params()
    for it in range(30):
        ncol = 3000
        rows = synth.alignment_window(ncol, seed=7000 + it, n_rows=4, snp=0.05, gap_rate=0.004, gap_mean=3)
        w = (rng.random(4) * (10.0 ** rng.integers(-3, 4))).astype(np.float32)       # weights make every score a 24-bit float
        c1, s1, m1, _, _ = _oracle.anchor_cols(rows, 2, weights=w)
        c2, s2, m2 = _emu_cols(rows, 2, w)
        assert np.array_equal(c1, c2) and _same_floats(s1, s2) and _same_floats(m1, m2), it


def test_kernel_value_path_equals_oracle_on_random_windows():
    rng = np.random.default_rng(5)
    for it in range(60):
        ncol = int(rng.integers(1, 9000))
        n1, n2 = (1, 1) if it % 3 else (int(rng.integers(1, 4)), int(rng.integers(1, 4)))
        rows = synth.alignment_window(ncol, seed=1000 + it, n_rows=n1 + n2, snp=float(rng.choice([0.02, 0.1, 0.3])),
                                      gap_rate=float(rng.choice([0.002, 0.01, 0.06])), gap_mean=int(rng.choice([2, 12, 150])),
                                      both_gap=float(rng.choice([0.0, 0.002, 0.03])))
        w = rng.random(n1 + n2).astype(np.float32) if it % 3 == 0 else np.ones(n1 + n2, dtype=np.float32)
        c1, s1, m1, _, _ = _oracle.anchor_cols(rows, n1, weights=w)
        c2, s2, m2 = _emu_cols(rows, n1, w)
        assert np.array_equal(c1, c2), it
        assert _same_floats(s1, s2) and _same_floats(m1, m2), it


@pytest.mark.skipif(not _oracle.have_ref_full(), reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_equals_reference_on_random_windows():
    rng = np.random.default_rng(9)
    total = 0
    for it in range(40):
        ncol = int(rng.integers(1, 6000))
        n1, n2 = (1, 1) if it % 4 else (int(rng.integers(1, 4)), int(rng.integers(1, 4)))
        rows = synth.alignment_window(ncol, seed=2000 + it, n_rows=n1 + n2, snp=float(rng.choice([0.02, 0.1, 0.3])),
                                      gap_rate=float(rng.choice([0.002, 0.01, 0.06])), gap_mean=int(rng.choice([2, 12, 150])),
                                      both_gap=float(rng.choice([0.0, 0.002, 0.03])))
        cols, score, smooth, w, fixed = _oracle.anchor_cols(rows, n1, use_ref=True)
        c2, s2, m2, _, _ = _oracle.anchor_cols(fixed, n1, weights=w)
        assert np.array_equal(cols, c2), it
        assert _same_floats(score, s2) and _same_floats(smooth, m2), it
        total += cols.size
    assert total > 200


def test_kernel_value_path_with_other_settings():
    """an extension penalty (every gap run is then a float sum in column order: the owner thread's loop, whatever the run's length),
    a smoothing ceiling that bites, other thresholds, window and spacing"""
    p = _oracle.anchor_default_params()
    p.smooth_ceil, p.min_best_col, p.min_smooth, p.smooth_window, p.anchor_spacing, p.gap_extend = 120.0, 100.0, 60.0, 7, 32, -5.0
    for it in range(12):
        rows = synth.alignment_window(2500, seed=8000 + it, gap_rate=0.01, gap_mean=int([3, 40, 200][it % 3]))
        c1, s1, m1, _, _ = _oracle.anchor_cols(rows, 1, params=p)
        c2, s2, m2 = _emu_cols(rows, 1, np.ones(2, dtype=np.float32), p)
        assert np.array_equal(c1, c2) and _same_floats(s1, s2) and _same_floats(m1, m2), it


def test_kernel_value_path_at_the_edges_of_the_exactness_tests():
    """weights that push the scores to 1e8 and down to 1e-4 (sums that do / do not fit 24 bits, fractions of every size), rows of
    wildcards only, lengths around the smoothing window and around the tile size of the kernel: the speculation in the smoothing
    chain may take or refuse what it likes, the answer has to be the oracle's"""
    rng = np.random.default_rng(23)
    cases = []
    for scale in (1e6, 3.0e5, 1024.0, 0.5, 1.0 / 3.0, 1e-4):
        rows = synth.alignment_window(2600, seed=int(scale * 7) % 1000 + 1, snp=0.05, gap_rate=0.003)
        cases.append((rows, 1, np.array([scale, 1.0], dtype=np.float32)))
        cases.append((rows, 1, np.array([scale, scale], dtype=np.float32)))
    for ncol in (21, 22, 23, 42, 43, 511 + 21, 512 + 21, 513 + 21, 1023 + 21, 1024 + 21, 1025 + 21, 2048 + 21, 2049 + 21):
        cases.append((synth.alignment_window(ncol, seed=ncol, gap_rate=0.004), 1, np.ones(2, dtype=np.float32)))
    wild = np.full((2, 900), ord("N"), dtype=np.uint8)
    cases.append((wild, 1, np.ones(2, dtype=np.float32)))
    half = synth.alignment_window(900, seed=3)
    half[1, :450] = ord("N")
    cases.append((half, 1, np.ones(2, dtype=np.float32)))
    for rows, n1, w in cases:
        c1, s1, m1, _, _ = _oracle.anchor_cols(rows, n1, weights=w)
        c2, s2, m2 = _emu_cols(rows, n1, w)
        assert np.array_equal(c1, c2) and _same_floats(s1, s2) and _same_floats(m1, m2), (rows.shape, w)
