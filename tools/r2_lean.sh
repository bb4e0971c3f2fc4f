#!/bin/bash
# lean A/B session: parity of the enumeration on small inputs, the three grouping variants on the headline workload, one ncu capture
set -u
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
O=gpurun_out/${TAG:-lean}
mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider --timeout 120 -k "mums or chunked or find_mums_into or sharded" > $O/pytest_quick.log 2>&1
echo "rc=$?" >> $O/pytest_quick.log
SHORT="--steps 10 --warmup 3 --no-dp --no-cpu --no-buildindex --no-sml"
for v in ${VARIANTS:-0 1 2}; do
  MAUVE_CUDA_GROUP_VARIANT=$v timeout 300 python bench.py $SHORT > $O/bench_var$v.json 2> $O/bench_var$v.err
done
PROF="python bench.py --steps 2 --warmup 3 --no-dp --no-cpu --no-buildindex --no-sml"
MAUVE_CUDA_GROUP_VARIANT=${NCU_VARIANT:-0} timeout 600 ncu --set full --clock-control none --import-source on -k regex:'bk_group3_kernel' -s 3 -c 1 -o $O/group3 -f $PROF > /dev/null 2> $O/group3.err
if [ -n "${HMM:-}" ]; then
  timeout 300 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 200 -k "hmm" > $O/pytest_hmm.log 2>&1
  echo "rc=$?" >> $O/pytest_hmm.log
  timeout 300 python tools/hmm_time.py > $O/hmm_time.log 2>&1
fi
echo done
