#!/usr/bin/env python
"""CPU check of the 'fused cell update' recipe in experiments/README.md: the recurrence of csrc/dp.cu (header comment) evaluated cell by
cell as dp.cu does today (D' = D - 200, I' = I - 200, M - 400 carried) and in the biased form of the recipe (D'' = D' + 400, I'' = I' + 400,
unbiased M carried, best = max(max(D'', I'') - 400, M)) gives the same four traceback predicates in every cell and the same final
(M, D', I') triple.  Plain Python over small random regions (ties are frequent: scores are multiples of small integers)."""
import numpy as np

NINF = -(1 << 29)
SUB = np.array([[151, -54, 29, -63], [-54, 160, -65, 29], [29, -65, 160, -54], [-63, 29, -54, 151]])  # NUC_SP + 60 (MU/nucmx.cpp:8-25), A C G T


def run(a, b, biased):
    la, lb = len(a), len(b)
    bias = 400 if biased else 0
    # row 0 / column 0 as in dp.cu: best[0][0] = 0 (-200 if la == 1), best[i][0] = best[0][j] = -200; no M, D', I' outside the matrix
    best_prev = [(-200 if la == 1 else 0)] + [-200] * lb
    Mrow_prev = [NINF] * (lb + 1)      # the CARRIED M of row i-1: M - 400 today, M in the biased form; NW_NINF outside the matrix in both
    Drow_prev = [NINF] * (lb + 1)      # D state of row i-1 (biased or not)
    bits = np.zeros((la, lb), dtype=np.uint8)
    last = None
    for i in range(la):
        best_row = [-200] + [0] * lb
        Mrow = [NINF] * (lb + 1)
        Drow = [NINF] * (lb + 1)
        I_left, M_left = NINF, NINF
        for j in range(1, lb + 1):
            M = SUB[a[i], b[j - 1]] + best_prev[j - 1]
            upM, upD = Mrow_prev[j], Drow_prev[j]
            D = max(upD, upM)
            I = max(I_left, M_left)
            if biased:
                best = max(max(D, I) - 400, M)
                b0 = best > M
                carried = M
            else:
                best = max(M, max(D, I))
                b0 = max(D, I) > M
                carried = M - 400
            b1, b2, b3 = I > D, upM >= upD, M_left >= I_left
            bits[i, j - 1] = b0 | (b1 << 1) | (b2 << 2) | (b3 << 3)
            Mrow[j], Drow[j], best_row[j] = carried, D, best
            I_left, M_left = I, carried
            last = (M, D + 200 - bias, I + 200 - bias)   # what nw_region stores as the result: (M, D, I) = (M, D' + 200, I' + 200)
        best_prev, Mrow_prev, Drow_prev = best_row, Mrow, Drow
    return bits, last


try:   # the oracle is the checker here as in tests/
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
    import _oracle
    orc = _oracle.oracle_checker()
except Exception:
    orc = None
rng = np.random.default_rng(5)
cells = 0
for t in range(300):
    la, lb = int(rng.integers(1, 60)), int(rng.integers(1, 60))
    a = rng.integers(0, 4, la)
    b = a[:lb].copy() if (t % 3 == 0 and lb <= la) else rng.integers(0, 4, lb)   # near-identical pairs: long diagonals, many ties
    x, lx = run(a, b, False)
    y, ly = run(a, b, True)
    # comparisons against NW_NINF (first row / column) agree too: both forms start the carried values at NW_NINF, as dp.cu does
    assert np.array_equal(x, y), (t, la, lb)
    assert lx == ly or min(lx) < NINF // 2, (lx, ly)
    if orc is not None and t < 60:   # anchor: the reference's score (oracle restatement of NWSmall) = max of the final triple
        edges, osc = orc.nw_align(bytes(b"ACGT"[k] for k in a), bytes(b"ACGT"[k] for k in b))
        assert max(ly) == osc, (t, ly, osc)
    cells += la * lb
print("biased and unbiased recurrences agree on every traceback predicate: 300 regions, %d cells%s" % (cells, "; scores of 60 regions equal the oracle's" if orc else ""))
