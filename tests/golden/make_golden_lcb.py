"""Mints tests/golden/lcb.npz by running the REFERENCE's own EliminateOverlaps_v2, LengthFilter, IdentifyBreakpoints and
ComputeLCBs_v2 (oracle/_ref/libmauve_ref_full.so = unmodified /root/reference sources, oracle/ref_driver_full.cpp).

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_lcb.py
  mds42_*        BASELINE config 1: the golden 29,403-row match list of the MDS42 pair through
                 EliminateOverlaps_v2(ml) [+ LengthFilter(MIN_ANCHOR_LENGTH + 3 = 12)]  (pairwiseAnchorSearch, LM/ProgressiveAligner.cpp:656-660)
                 EliminateOverlaps_v2(ml, true), IdentifyBreakpoints, ComputeLCBs_v2    (pairwise LCB set-up, :3408-3418)
                 (the genome-1 ordering of the second pass meets 25 ties there: the result pins libstdc++'s introsort)
  syn_*          the match list of a synthetic 300 kbp pair with inversions and a repeat family (rows stored), same calls
  rand_*         hand-made lists (rows stored): dense overlaps on both strands, ties, nested matches, single rows
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _oracle  # noqa: E402
from mauve_py_b200 import synth  # noqa: E402

MIN_LEN = 12   # MIN_ANCHOR_LENGTH + 3, LM/Aligner.h:266, LM/ProgressiveAligner.cpp:660


def random_lists():
    rng = np.random.default_rng(20261022)
    out = []
    for it in range(40):
        n = int(rng.integers(1, 500))
        G = int(rng.integers(2000, 200000))
        s0 = rng.integers(1, G, n)
        ln = rng.integers(5, 400, n)
        s1 = rng.integers(1, G, n) * rng.choice([1, -1], n)
        k = n // 2
        s0[:k] = np.sort(rng.integers(1, max(G // 10, 2), k))
        s1[:k] = s0[:k] + rng.integers(-3, 4, k)
        s1[s1 == 0] = 1
        out.append(np.stack([ln, s0, s1], 1).astype(np.int64))
    out.append(np.array([[50, 10, 10]], dtype=np.int64))
    out.append(np.array([[50, 10, 10], [50, 10, -200], [20, 30, 30], [100, 5, 400]], dtype=np.int64))   # equal starts, nested
    return out


def run_all(rows):
    d = {}
    d["elim0"] = _oracle.eliminate_overlaps(rows, False, 0, use_ref=True)[0]
    d["elim0_min"] = _oracle.eliminate_overlaps(rows, False, MIN_LEN, use_ref=True)[0]
    d["elim1"] = _oracle.eliminate_overlaps(rows, True, 0, use_ref=True)[0]
    so, bp, _ = _oracle.lcbs(d["elim1"], use_ref=True) if d["elim1"].shape[0] else (np.zeros((0, 3), np.int64), np.zeros(0, np.uint64), None)
    d["lcb_sorted"], d["lcb_bp"] = so, bp
    return d


def main():
    out = {}
    rows = np.load(os.path.join(HERE, "mums_mds42.npz"))["rows_w15_r3"]
    for k, v in run_all(rows).items():
        out["mds42_" + k] = v
    a, b = synth.small_pair(300000, seed=41, snp=0.02, n_inv=4)
    unit = synth.random_genome(700, 0.5, synth.rng_for(5)).tobytes()
    a = a[:100000] + unit + a[100000:200000] + unit + a[200000:]
    b = b[:50000] + unit + b[50000:250000] + unit + b[250000:]
    chk = _oracle.ref_checker()
    seed = chk.get_seed(11, 0)
    syn, _ = chk.find_mums(a, b, seed, 0)
    out["syn_rows"] = np.ascontiguousarray(syn, dtype=np.int64)
    for k, v in run_all(out["syn_rows"]).items():
        out["syn_" + k] = v
    lists = random_lists()
    out["rand_count"] = np.array(len(lists))
    for i, r in enumerate(lists):
        out["rand%d_rows" % i] = r
        for k, v in run_all(r).items():
            out["rand%d_%s" % (i, k)] = v
    np.savez_compressed(os.path.join(HERE, "lcb.npz"), **out)
    print("mds42:", out["mds42_elim0"].shape, out["mds42_elim0_min"].shape, out["mds42_elim1"].shape, "LCBs", out["mds42_lcb_bp"].size)
    print("syn:", out["syn_rows"].shape, out["syn_elim1"].shape, "LCBs", out["syn_lcb_bp"].size)


if __name__ == "__main__":
    main()
