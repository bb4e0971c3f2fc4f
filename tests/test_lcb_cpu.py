"""SURVEY.md 8f-1 on the CPU: the C restatement of EliminateOverlaps_v2 / LengthFilter / IdentifyBreakpoints / ComputeLCBs_v2
(oracle/mauve_oracle.c, incl. its restatement of libstdc++'s std::sort, which decides where tied rows end up) against the golden
vectors the reference's own functions produced (tests/golden/lcb.npz, make_golden_lcb.py) and, where oracle/_ref is present,
against those functions directly on fresh random lists."""
import numpy as np
import pytest

import _golden
import _oracle

MIN_LEN = 12


def _cases(z):
    names = ["mds42", "syn"] + ["rand%d" % i for i in range(int(z["rand_count"]))]
    for nm in names:
        rows = _golden.npz("mums_mds42.npz")["rows_w15_r3"] if nm == "mds42" else z[nm + "_rows"]
        yield nm, np.ascontiguousarray(rows, dtype=np.int64)


def test_oracle_equals_the_reference_goldens():
    z = _golden.npz("lcb.npz")
    ties_seen = 0
    for nm, rows in _cases(z):
        for key, both, ml in (("elim0", False, 0), ("elim0_min", False, MIN_LEN), ("elim1", True, 0)):
            got, ties = _oracle.eliminate_overlaps(rows, both, ml)
            assert np.array_equal(got, z["%s_%s" % (nm, key)]), (nm, key)
            ties_seen += ties
        e1 = z[nm + "_elim1"]
        if e1.shape[0]:
            so, bp, _ = _oracle.lcbs(e1)
            assert np.array_equal(so, z[nm + "_lcb_sorted"]) and np.array_equal(bp, z[nm + "_lcb_bp"]), nm
            assert bp[-1] == e1.shape[0] - 1 and np.all(np.diff(bp.astype(np.int64)) > 0)
    assert ties_seen > 0   # the goldens do exercise the order of tied rows (MDS42: 25 ties in the second pass)
    assert z["mds42_elim0"].shape[0] == 27008 and z["mds42_elim1"].shape[0] == 27161 and z["mds42_lcb_bp"].size == 2697


def test_tied_rows_follow_introsort_not_a_stable_sort():
    """the MDS42 list with eliminate_both: ordering the second pass with a stable sort instead gives a different list -- which is
    why the restatement (and csrc/lcb.cu) run libstdc++'s algorithm when keys tie"""
    z = _golden.npz("lcb.npz")
    rows = _golden.npz("mums_mds42.npz")["rows_w15_r3"]
    got, ties = _oracle.eliminate_overlaps(rows, True, 0)
    assert ties == 25 and got.shape[0] == 27161
    assert np.array_equal(got, z["mds42_elim1"])


@pytest.mark.skipif(not _oracle.have_ref_full(), reason="oracle/_ref not built here")
def test_oracle_equals_the_reference_on_fresh_lists():
    rng = np.random.default_rng(77)
    for it in range(150):
        n = int(rng.integers(1, 400))
        G = int(rng.integers(2000, 200000))
        s0 = rng.integers(1, G, n)
        ln = rng.integers(5, 400, n)
        s1 = rng.integers(1, G, n) * rng.choice([1, -1], n)
        k = n // 2
        s0[:k] = np.sort(rng.integers(1, max(G // 10, 2), k))
        s1[:k] = s0[:k] + rng.integers(-3, 4, k)
        s1[s1 == 0] = 1
        rows = np.stack([ln, s0, s1], 1).astype(np.int64)
        for both in (False, True):
            for ml in (0, MIN_LEN):
                a, _ = _oracle.eliminate_overlaps(rows, both, ml, use_ref=True)
                b, _ = _oracle.eliminate_overlaps(rows, both, ml)
                assert np.array_equal(a, b), (it, both, ml)
        e1, _ = _oracle.eliminate_overlaps(rows, True, 0)
        if e1.shape[0]:
            so, bp, _ = _oracle.lcbs(e1, use_ref=True)
            so2, bp2, _ = _oracle.lcbs(e1)
            assert np.array_equal(so, so2) and np.array_equal(bp, bp2), it
