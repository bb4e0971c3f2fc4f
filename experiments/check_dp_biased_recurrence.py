#!/usr/bin/env python
"""CPU check of the 'fused cell update' recipe in experiments/README.md: the recurrence of csrc/dp.cu (header comment) evaluated cell by
cell as dp.cu does today (D' = D - 200, I' = I - 200, M - 400 carried) and in the biased form of the recipe (D'' = D' + 400, I'' = I' + 400,
unbiased M carried, best = max(max(D'', I'') - 400, M)) gives the same four traceback predicates in every cell and the same final
(M, D', I') triple.  Plain Python over small random regions (ties are frequent: scores are multiples of small integers)."""
import numpy as np

import os
import sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from _properties import NW_NINF as NINF, nw_integer_recurrence as run  # noqa: E402

try:   # the oracle is the checker here as in tests/
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
