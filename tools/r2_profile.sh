#!/bin/bash
# Round-2 profile session (one B200): the whole -m gpu suite, the full bench line, then the evidence for profiles/: the launch list of a
# bench step and one `ncu --set full` capture of every kernel the numbers rest on.  Nothing printed under ncu is a bench value.
set -u
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
O=gpurun_out/${TAG:-prof}
mkdir -p $O
export PYTHONUNBUFFERED=1
nvidia-smi -L > $O/gpus.txt 2>&1
if [ -z "${SKIP_TESTS:-}" ]; then
  timeout 1700 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 --durations=12 > $O/pytest_gpu.log 2>&1
  echo "rc=$?" >> $O/pytest_gpu.log
fi
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench_full.json 2> $O/bench_full.err
echo "rc=$?" >> $O/bench_full.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
PROF="python bench.py --steps 2 --warmup 3 --no-dp --no-cpu --no-buildindex --no-sml"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv $PROF > /dev/null 2> $O/launches.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'bk_group3_kernel|bkf_scatter1_kernel|bkf_scatter2_kernel' -s 9 -c 3 \
    -o $O/partition -f $PROF > /dev/null 2> $O/partition.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'candidate_kernel|extend_solid_kernel|order_ties_kernel|pack_kernel' -s 12 -c 4 \
    -o $O/finish -f $PROF > /dev/null 2> $O/finish.err
cat > $O/_sml_probe.py <<'PY'
import sys, ctypes as C
sys.path.insert(0, ".")
import numpy as np
import mauve_py_b200 as mp
from mauve_py_b200 import synth
from mauve_py_b200._capi import check
check(mp.lib().mcu_init(0))
g = synth.random_genome(100_000_000, 0.41, synth.rng_for(3))
n = C.c_uint64(0)
for w, r in ((19, 3), (15, 3)):
    seed = mp.getSeed(w, r)
    for _ in range(2):
        check(mp.lib().mcu_sml_build(g.ctypes.data, g.size, seed, None, None, None, C.byref(n)))
    st = np.zeros(6, dtype=np.float32)
    mp.lib().mcu_sml_last_stats(st.ctypes.data)
    print(w, r, hex(seed), n.value, st.tolist())
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'rs_onesweep_kernel|seedgen_kernel' -s 6 -c 3 -o $O/sml -f python $O/_sml_probe.py > $O/sml.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'nw_forward_kernel' -c 2 -o $O/dp -f python tools/config5_dp.py --regions 1500 --chunk 1500 > $O/dp_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'hmm_exact_chain_warp_kernel' -c 1 -o $O/hmm_warp -f python tools/hmm_time.py --one 500000 > /dev/null 2> $O/hmm_ncu.err
timeout 300 python tools/hmm_time.py > $O/hmm_time.log 2>&1
# the anchor-column kernel (8f-4): ONE 20,000-column window per call -- the aligner's case -- captured, and its phase clocks on three kinds of window
cat > $O/_cols_probe.py <<'PY'
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
    print(name, "anchor columns", c.size, "segments exact/chain", k[:2].tolist(), "SM cycles scoring/smoothing/best/ends/walk/picks", k[2:].tolist())
PY
timeout 300 python $O/_cols_probe.py > $O/cols_phases.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'anchor_cols_kernel' -s 2 -c 1 -o $O/anchor_cols -f python $O/_cols_probe.py > $O/cols_ncu.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $O/nvidia_smi.csv 2>&1
echo done
