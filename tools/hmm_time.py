import sys, time
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mauve_py_b200 as mp
from mauve_py_b200 import synth
from mauve_py_b200._capi import check
check(mp.lib().mcu_init(0))
params = mp.libmems.hmm_params(0.5, 1e-5, 1e-9, 0.7)
def run(name, seqs):
    mp.run_batch(seqs, params, True)
    t=time.time(); _,_,ms = mp.run_batch(seqs, params, True); w=time.time()-t
    tot=sum(len(s) for s in seqs)
    print(name, len(seqs), tot, "device_ms %.3f wall_ms %.1f  Gcol/s %.3f" % (ms, w*1e3, tot/ms/1e6))
def counters():
    c = np.zeros(3, dtype=np.uint64)
    mp.lib().mcu_test_hmm_counters(c.ctypes.data)
    return "columns %d rounds %d fp64 columns %d" % tuple(int(x) for x in c)
if "--one" in sys.argv:
    n = int(sys.argv[sys.argv.index("--one") + 1])
    run("one", [synth.hmm_string(n, seed=1, block=3000)])
    sys.exit(0)
one = synth.hmm_string(5_000_000, seed=1, block=3000)
run("one5M", [one])
print(counters())
run("one500k", [one[:500000]])
pairs = synth.dp_pairs(512, 100, 10000, seed=20261020)
run("bench512", [synth.hmm_string(len(p[0]), seed=i, block=300) for i, p in enumerate(pairs)])
base = synth.hmm_string(3000, seed=3, block=300)
run("many100k", [base]*100000)
