"""One rank's share of a sharded step on one GPU (for ncu): python tools/shard_step.py <mbp> <shard> <nshard> [steps]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mauve_py_b200 as mp
from mauve_py_b200 import synth
from mauve_py_b200._capi import check

mbp, shard, nshard = float(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
check(mp.lib().mcu_init(0))
a, b = synth.config3_pair(n=int(mbp * 1e6))
seed = mp.getSeed(mp.getDefaultSeedWeight((a.size + b.size) // 2), mp.CODING_SEED)
s = mp.AnchorSession()
s.upload(a.tobytes(), b.tobytes())
for _ in range(steps):
    n = s.run(seed, shard, nshard)
    print(n, [round(float(x), 3) for x in s.stage_ms[8:15]], "total %.3f" % s.stage_ms[6])
