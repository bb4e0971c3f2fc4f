#!/bin/bash
# quick A/B of the sorted-mer-list build (seed generation + onesweep radix sort) on a 100 Mbp genome: device times from mcu_sml_last_stats
OUT=gpurun_out/${TAG:-smlq}
mkdir -p $OUT
cat > $OUT/_sml.py <<'PY'
import sys, ctypes as C
sys.path.insert(0, ".")
import numpy as np
import mauve_py_b200 as mp
from mauve_py_b200 import synth
from mauve_py_b200._capi import check
check(mp.lib().mcu_init(0))
g = synth.random_genome(100_000_000, 0.41, synth.rng_for(3))
n = C.c_uint64(0)
for w, r in ((19, 3), (15, 3), (11, 0), (21, 0)):
    seed = mp.getSeed(w, r)
    best = None
    for _ in range(4):
        check(mp.lib().mcu_sml_build(g.ctypes.data, g.size, seed, None, None, None, C.byref(n)))
        st = np.zeros(6, dtype=np.float32)
        mp.lib().mcu_sml_last_stats(st.ctypes.data)
        if best is None or st[2] < best[2]:
            best = st.copy()
    per_pass = best[2] / best[3]
    gbs = 2 * (best[4] + 4) * best[5] / (per_pass * 1e-3) / 1e9
    print("weight %d rank %d: seedgen %.3f ms, sort %.3f ms in %d passes of %d-byte keys = %.3f ms per pass = %.0f GB/s" % (w, r, best[1], best[2], int(best[3]), int(best[4]), per_pass, gbs))
PY
python $OUT/_sml.py > $OUT/sml.log 2>&1
cat $OUT/sml.log
if [ -n "${WITH_TESTS:-}" ]; then
  timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "sml or sort" > $OUT/pytest_sml.log 2>&1; tail -3 $OUT/pytest_sml.log
fi
