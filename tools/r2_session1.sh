#!/bin/bash
# Round-2 GPU session 1 (one B200): parity of the new paths first, then A/B numbers, then profiles.  Everything under its own timeout.
set -u
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
O=gpurun_out/s1
mkdir -p $O
export PYTHONUNBUFFERED=1
nvidia-smi -L > $O/gpus.txt 2>&1

# 1. the new grouping kernel + chunked upload on small inputs first (a hang here must not eat the session)
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider --timeout 120 -k "mums_vs_oracle or chunked or find_mums_into or sharded_entry" > $O/pytest_quick.log 2>&1
echo "rc=$?" >> $O/pytest_quick.log
# 2. whole parity file + full-size configs (golden sha1 of the 100 Mbp list)
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_zz_fullsize_gpu.py -m gpu -q -p no:cacheprovider --timeout 300 --durations=15 > $O/pytest_parity.log 2>&1
echo "rc=$?" >> $O/pytest_parity.log

# 3. A/B of the grouping kernel and of the upload overlap on the headline workload
SHORT="--steps 10 --warmup 3 --no-dp --no-cpu --no-buildindex --no-sml"
timeout 300 python bench.py $SHORT > $O/bench_v2.json 2> $O/bench_v2.err
MAUVE_CUDA_GROUP_V1=1 timeout 300 python bench.py $SHORT > $O/bench_v1.json 2> $O/bench_v1.err
MAUVE_CUDA_NO_OVERLAP=1 timeout 300 python bench.py $SHORT > $O/bench_nooverlap.json 2> $O/bench_nooverlap.err

# 4. DP: fused cell update (experiments/r02_dp_fused_cell_update.patch, prebuilt variant)
V=build/variants/libmauve_cuda_dpfused.so
if [ -f $V ]; then
  MAUVE_CUDA_LIB=$PWD/$V timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider --timeout 200 -k "nw" > $O/pytest_dpfused.log 2>&1
  echo "rc=$?" >> $O/pytest_dpfused.log
  timeout 300 python tools/config5_dp.py --regions 6000 --chunk 6000 > $O/dp_base.json 2> $O/dp_base.err
  MAUVE_CUDA_LIB=$PWD/$V timeout 300 python tools/config5_dp.py --regions 6000 --chunk 6000 > $O/dp_fused.json 2> $O/dp_fused.err
fi

# 5. profiles: launch list of one step, full captures of the three partition / grouping kernels, of the SML build and of the DP
PROF="python bench.py --steps 2 --warmup 3 --no-dp --no-cpu --no-buildindex --no-sml"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv $PROF > /dev/null 2> $O/launches.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'bk_group2_kernel|bkf_scatter1_kernel|bkf_scatter2_kernel' -s 9 -c 3 \
    -o $O/partition -f $PROF > /dev/null 2> $O/partition.err
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

# 6. the whole bench line
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench_full.json 2> $O/bench_full.err
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $O/nvidia_smi.csv 2>&1
echo done
