#!/bin/bash
# GPU session for the anchor-column kernel (8f-4): parity tests, the seam inside the reference binary, a first timing
TAG=${TAG:-cols}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $OUT/gpus.txt 2>&1
timeout 600 python -m pytest tests/test_zzzz_next_rows_gpu.py -x -q -m gpu -k "anchor_cols" > $OUT/pytest_cols.log 2>&1
echo "pytest_cols rc=$?" >> $OUT/pytest_cols.log
tail -5 $OUT/pytest_cols.log
timeout 900 python -m pytest tests/test_zzz_buildindex.py -x -q -m gpu -k "dropin_mds42" -s > $OUT/pytest_seam.log 2>&1
echo "pytest_seam rc=$?" >> $OUT/pytest_seam.log
tail -8 $OUT/pytest_seam.log
timeout 600 python - > $OUT/time_cols.log 2>&1 <<'PY'
import sys, json, argparse
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import bench, mauve_py_b200 as mp
from mauve_py_b200 import synth
from mauve_py_b200._capi import check
check(mp.lib().mcu_init(0))
for nw in (1, 148, 1184, 4736):
    a = argparse.Namespace(anchor_windows=nw, no_cpu=(nw != 1184))
    print(json.dumps(bench.measure_anchor_cols(mp, synth, a)))
PY
tail -6 $OUT/time_cols.log
timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; tail -2 $OUT/smoke.log
cat > $OUT/_one.py <<'PY'
import sys
sys.path.insert(0, ".")
import numpy as np
import mauve_py_b200 as mp
from mauve_py_b200 import synth
from mauve_py_b200._capi import check
check(mp.lib().mcu_init(0))
clean = len(sys.argv) > 1 and sys.argv[1] == "clean"
w = synth.alignment_window(20000, seed=900, **(dict(snp=0.02, gap_rate=0.0005, diverged_blocks=False) if clean else {}))
for _ in range(3):
    c = mp.FindAnchorColsPP(w[:1], w[1:])
k = np.zeros(8, dtype=np.uint64)
mp.lib().mcu_test_anchor_counters(k.ctypes.data)
print(c.size, k.tolist())
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'anchor_cols_kernel' -s 2 -c 1 -o $OUT/one_window -f python $OUT/_one.py > $OUT/ncu_one.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'anchor_cols_kernel' -s 2 -c 1 -o $OUT/one_window_clean -f python $OUT/_one.py clean > $OUT/ncu_one_clean.log 2>&1
tail -2 $OUT/ncu_one.log
