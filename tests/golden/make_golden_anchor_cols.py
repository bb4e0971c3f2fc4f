"""Mints tests/golden/anchor_cols.npz by running the REFERENCE's own FindAnchorColsPP, LetterObjScoreXP and WindowSmooth
(oracle/_ref/libmauve_ref_full.so = unmodified /root/reference sources, oracle/ref_driver_full.cpp: ref_anchor_cols) after the set-up
MuscleInterface::ProfileAlignFast does (LM/MuscleInterface.cpp:1086-1106), the rows' weights from PrepareMSAforScoring as in
AnchoredProfileProfile (MU/anchoredpp.cpp:454-455).

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_anchor_cols.py
Per window k:  w{k}_rows (characters as scored, i.e. after MSA::FixAlpha), w{k}_n1 (rows of the first alignment), w{k}_weights,
               w{k}_cols (anchor columns), w{k}_score / w{k}_smooth (per-column scores, float32 bit patterns kept as they are)
and `settings` / `letters`: the score matrix, gap penalties, thresholds and the character table the reference had in force.
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


def windows():
    """(rows, n1) per window: the two-genome form of the pipeline at many lengths and gap densities, then hand-made edge cases,
    then windows of two alignments with more rows (weights from the guide tree)"""
    out = []
    for k, ncol in enumerate([1, 2, 20, 21, 22, 23, 43, 96, 97, 200, 1000, 3000, 8000, 20000]):
        out.append((synth.alignment_window(ncol, seed=100 + k), 1))
    out.append((synth.alignment_window(6000, seed=201, snp=0.02, gap_rate=0.002, diverged_blocks=False), 1))     # nearly every column is "best"
    out.append((synth.alignment_window(6000, seed=202, snp=0.5, gap_rate=0.05), 1))                              # hardly any
    out.append((synth.alignment_window(4000, seed=203, gap_rate=0.08, gap_mean=3, both_gap=0.02), 1))            # gap runs touching each other
    out.append((synth.alignment_window(4000, seed=204, gap_rate=0.004, gap_mean=300), 1))                        # long runs
    a = np.frombuffer(b"ACGTACGTACGTACGTACGTACGTACGTACGTACGTACGT", dtype=np.uint8)
    def two(x, y):
        return np.stack([np.frombuffer(x, dtype=np.uint8), np.frombuffer(y, dtype=np.uint8)])
    out.append((two(b"-" * 40, b"-" * 40), 1))                                                # nothing but gap columns
    out.append((two(b"-" * 10 + bytes(a[:30]), bytes(a[:30]) + b"-" * 10), 1))                # terminal gaps on both sides
    out.append((two(b"--" + bytes(a[:36]) + b"--", b"--" + bytes(a[:36]) + b"--"), 1))        # all-gap columns at both ends
    out.append((two(bytes(a[:10]) + b"-----" + bytes(a[:25]), bytes(a[:15]) + b"-----" + bytes(a[:20])), 1))   # a gap in row 1 followed at once by one in row 2
    out.append((two(bytes(a[:10]) + b"---" + b"AAA" + bytes(a[:24]), bytes(a[:10]) + b"AAA" + b"---" + bytes(a[:24])), 1))
    out.append((two(bytes(a[:10]) + b"--..-" + bytes(a[:25]), bytes(a[:12]) + b"-" + bytes(a[:27])), 1))       # '.' is a gap too; one row-1 gap column, then both
    out.append((two(b"ACGTNNNNacgtRYKM" + bytes(a[:24]), b"ACGTACGTACGTACGT" + bytes(a[:24])), 1))             # wildcards, lower case
    out.append((two(b"ACGTJZ*?ACGTACGT" + bytes(a[:24]), b"ACGTACGTACGTACGT" + bytes(a[:24])), 1))             # characters FixAlpha rewrites
    for k, (r1, r2, ncol) in enumerate([(2, 1, 500), (1, 3, 800), (2, 2, 1500), (3, 3, 2500), (4, 2, 300)]):
        out.append((synth.alignment_window(ncol, seed=300 + k, n_rows=r1 + r2, snp=0.08), r1))
    return out


def main():
    out = {}
    settings, letters = _oracle.anchor_settings_ref()
    out["settings"] = settings
    out["letters"] = letters
    ws = windows()
    total = 0
    for k, (rows, n1) in enumerate(ws):
        cols, score, smooth, weights, fixed = _oracle.anchor_cols(rows, n1, use_ref=True)
        out["w%d_rows" % k] = fixed
        out["w%d_n1" % k] = np.int64(n1)
        out["w%d_weights" % k] = weights
        out["w%d_cols" % k] = cols
        out["w%d_score" % k] = score
        out["w%d_smooth" % k] = smooth
        total += rows.shape[1]
        # the restatement, on the spot
        c2, s2, m2, _, _ = _oracle.anchor_cols(fixed, n1, weights=weights)
        ok = np.array_equal(cols, c2) and np.array_equal(score.view(np.uint32), s2.view(np.uint32)) and np.array_equal(smooth.view(np.uint32), m2.view(np.uint32))
        print("window %2d: %d+%d rows x %5d columns, %4d anchor columns, weights %s, oracle %s" % (k, n1, rows.shape[0] - n1, rows.shape[1], cols.size, np.round(weights, 4), "equal" if ok else "DIFFERENT"))
    out["n_windows"] = np.int64(len(ws))
    path = os.path.join(HERE, "anchor_cols.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", total, "columns")


if __name__ == "__main__":
    main()
