#!/bin/bash
# compute-sanitizer over the bucketed match-finding pipeline (TMA-fed grouping, partition passes, extension, order) on a 5 Mbp pair
OUT=gpurun_out/${TAG:-sanmums}
mkdir -p $OUT
cat > $OUT/_mums.py <<'PY'
import sys, hashlib
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import mauve_py_b200 as mp
from mauve_py_b200 import synth
from mauve_py_b200._capi import check
check(mp.lib().mcu_init(0))
a, b = synth.config2_pair(int(sys.argv[1]) if len(sys.argv) > 1 else 5_000_000)
seed = mp.getSeed(mp.getDefaultSeedWeight((len(a) + len(b)) // 2), 3)
rows, stats = mp.libmems.find_mums(a, b, seed)
print("rows", rows.shape[0], hashlib.sha1(np.ascontiguousarray(rows).tobytes()).hexdigest(), stats.tolist())
PY
python $OUT/_mums.py > $OUT/plain.log 2>&1; tail -1 $OUT/plain.log
for tool in memcheck racecheck; do
  timeout 1200 compute-sanitizer --tool $tool --print-limit 10 python $OUT/_mums.py > $OUT/mums_$tool.log 2>&1; tail -3 $OUT/mums_$tool.log
done
