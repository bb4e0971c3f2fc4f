#!/bin/bash
# quick look at the anchor-column kernel's phase clocks on one window (dirty / clean)
OUT=gpurun_out/${TAG:-colsq}
mkdir -p $OUT
cat > $OUT/_one.py <<'PY'
import sys
sys.path.insert(0, ".")
import numpy as np
import mauve_py_b200 as mp
from mauve_py_b200 import synth
from mauve_py_b200._capi import check
check(mp.lib().mcu_init(0))
for name, kw in (("bench", {}), ("clean", dict(snp=0.02, gap_rate=0.0005, diverged_blocks=False)), ("gappy", dict(gap_rate=0.05))):
    w = synth.alignment_window(20000, seed=900, **kw)
    for _ in range(3):
        c = mp.FindAnchorColsPP(w[:1], w[1:])
    k = np.zeros(8, dtype=np.uint64)
    mp.lib().mcu_test_anchor_counters(k.ctypes.data)
    print(name, c.size, k.tolist())
PY
python $OUT/_one.py > $OUT/one.log 2>&1
cat $OUT/one.log
