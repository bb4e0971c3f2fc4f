#!/bin/bash
# compute-sanitizer over the anchor-column kernel: memcheck and racecheck (shared-memory hazards of the two-buffer smoothing pipeline and
# of the block-wide scans), on a few windows in both launch forms
OUT=gpurun_out/${TAG:-san}
mkdir -p $OUT
cat > $OUT/_san.py <<'PY'
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import mauve_py_b200 as mp
import _oracle
from mauve_py_b200 import synth
from mauve_py_b200._capi import check
check(mp.lib().mcu_init(0))
small = [(synth.alignment_window(n, seed=40 + i, gap_rate=g), 1, None) for i, (n, g) in enumerate([(3000, 0.01), (1500, 0.0005), (700, 0.05), (40, 0.01), (2600, 0.002)])]
small.append((synth.alignment_window(900, seed=77, n_rows=4), 2, np.array([0.3, 0.7, 0.55, 0.45], dtype=np.float32)))
many = [(synth.alignment_window(300 + 7 * i, seed=500 + i), 1, None) for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 0)]
for wins in (small, many):
    if not wins:
        continue
    got = mp.libmems.FindAnchorColsPP_batch(wins)
    for (rows, n1, w), g in zip(wins, got):
        c = _oracle.anchor_cols(rows, n1, weights=w)[0]
        assert np.array_equal(c, g)
print("ok", len(small), len(many))
PY
timeout 600 compute-sanitizer --tool memcheck python $OUT/_san.py 320 > $OUT/memcheck.log 2>&1; tail -4 $OUT/memcheck.log
timeout 800 compute-sanitizer --tool racecheck --print-limit 20 python $OUT/_san.py 320 > $OUT/racecheck.log 2>&1; tail -4 $OUT/racecheck.log
timeout 600 compute-sanitizer --tool synccheck python $OUT/_san.py > $OUT/synccheck.log 2>&1; tail -4 $OUT/synccheck.log
