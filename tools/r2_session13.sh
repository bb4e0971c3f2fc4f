#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
O=gpurun_out/${TAG:-s13}
mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_zzzz_next_rows_gpu.py -m gpu -q -x -p no:cacheprovider --timeout 300 -k "eliminate" -s > $O/pytest_lcb.log 2>&1
echo "rc=$?" >> $O/pytest_lcb.log
timeout 900 python -m pytest tests/test_zzz_buildindex.py -m gpu -q -x -p no:cacheprovider --timeout 600 -k "seam_binaries_mds42 and sol" -s > $O/pytest_seam.log 2>&1
echo "rc=$?" >> $O/pytest_seam.log
python - > $O/lcb_time.log 2>&1 <<'PY'
import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import mauve_py_b200 as mp
import _golden
from mauve_py_b200._capi import check
check(mp.lib().mcu_init(0))
rows = _golden.npz("mums_mds42.npz")["rows_w15_r3"]
for both in (False, True):
    mp.EliminateOverlaps_v2(rows, both)
    t = time.perf_counter(); out, ties = mp.EliminateOverlaps_v2(rows, both, return_ties=True); dt = time.perf_counter() - t
    print("mds42 eliminate_both=%s: %d -> %d rows, ties %d, %.2f ms" % (both, rows.shape[0], out.shape[0], ties, 1e3 * dt))
t = time.perf_counter(); so, bp = mp.IdentifyBreakpoints(out); dt = time.perf_counter() - t
print("mds42 IdentifyBreakpoints: %d LCBs, %.2f ms" % (bp.size, 1e3 * dt))
PY
echo done
