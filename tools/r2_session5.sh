#!/bin/bash
# Round-2 GPU session 5 (one B200): split-genome final buckets + bk_group3 as a hash join with deferred write-out; the float
# wavefront kernel after its traceback fix.  Parity first, then A/B, profiles, the whole GPU suite.  Everything under its own timeout.
set -u
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
O=gpurun_out/s5
mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider --timeout 120 -k "mums or chunked or find_mums_into or sharded" > $O/pytest_quick.log 2>&1
echo "rc=$?" >> $O/pytest_quick.log
timeout 300 python -m pytest tests/test_zzzz_next_rows_gpu.py -m gpu -q -p no:cacheprovider --timeout 120 -k "nw_wild" > $O/pytest_wild.log 2>&1
echo "rc=$?" >> $O/pytest_wild.log
SHORT="--steps 10 --warmup 3 --no-dp --no-cpu --no-buildindex --no-sml"
for v in 0 1 2; do
  MAUVE_CUDA_GROUP_VARIANT=$v timeout 300 python bench.py $SHORT > $O/bench_var$v.json 2> $O/bench_var$v.err
done
MAUVE_CUDA_GROUP_V1=1 timeout 300 python bench.py $SHORT > $O/bench_v1.json 2> $O/bench_v1.err
PROF="python bench.py --steps 2 --warmup 3 --no-dp --no-cpu --no-buildindex --no-sml"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'bk_group3_kernel|bkf_scatter2_kernel' -s 6 -c 2 -o $O/group3 -f $PROF > /dev/null 2> $O/group3.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv $PROF > /dev/null 2> $O/launches.err
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 --durations=10 > $O/pytest_all.log 2>&1
echo "rc=$?" >> $O/pytest_all.log
echo done
