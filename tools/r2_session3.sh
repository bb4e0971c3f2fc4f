#!/bin/bash
# Round-2 GPU session 3 (one B200): the new grouping kernel (bk_group3: TMA-fed, one CAS table per bucket) -- parity on small inputs
# first, then A/B against bk_group on the headline workload, the whole GPU suite, an ncu capture.  Everything under its own timeout.
set -u
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
O=gpurun_out/s3
mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider --timeout 120 -k "mums or chunked or find_mums_into or sharded" > $O/pytest_quick.log 2>&1
echo "rc=$?" >> $O/pytest_quick.log
SHORT="--steps 10 --warmup 3 --no-dp --no-cpu --no-buildindex --no-sml"
timeout 300 python bench.py $SHORT > $O/bench_g3_s12.json 2> $O/bench_g3_s12.err
MAUVE_CUDA_GROUP_S11=1 timeout 300 python bench.py $SHORT > $O/bench_g3_s11.json 2> $O/bench_g3_s11.err
MAUVE_CUDA_GROUP_V1=1 timeout 300 python bench.py $SHORT > $O/bench_v1.json 2> $O/bench_v1.err
PROF="python bench.py --steps 2 --warmup 3 --no-dp --no-cpu --no-buildindex --no-sml"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'bk_group3_kernel' -s 3 -c 1 -o $O/group3 -f $PROF > /dev/null 2> $O/group3.err
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 --durations=15 > $O/pytest_all.log 2>&1
echo "rc=$?" >> $O/pytest_all.log
echo done
