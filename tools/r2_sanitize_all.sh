#!/bin/bash
# compute-sanitizer over one small pass of every part of the path (smoke(): SML build, match finding, DP, HMM, overlaps / LCBs, anchor
# columns) and over the long-string HMM kernel (warp chain with its verifier warps)
OUT=gpurun_out/${TAG:-sanall}
mkdir -p $OUT
cat > $OUT/_hmm.py <<'PY'
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import mauve_py_b200 as mp
import _oracle
from mauve_py_b200 import synth
from mauve_py_b200._capi import check
check(mp.lib().mcu_init(0))
params = mp.libmems.hmm_params(0.5, 1e-5, 1e-9, 0.7)
sym = synth.hmm_string(150000, seed=5, block=300)
pred, post = mp.run(sym, params, want_posterior=True)
opred, opost = _oracle.oracle_checker().hmm_run(sym, params)
assert pred == opred and np.array_equal(post.view(np.uint64), np.asarray(opost).view(np.uint64))
print("hmm ok")
PY
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 10 python __graft_entry__.py smoke > $OUT/smoke_$tool.log 2>&1; tail -3 $OUT/smoke_$tool.log
  timeout 900 compute-sanitizer --tool $tool --print-limit 10 python $OUT/_hmm.py > $OUT/hmm_$tool.log 2>&1; tail -3 $OUT/hmm_$tool.log
done
