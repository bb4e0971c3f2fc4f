#!/bin/bash
# HMM session (one B200): the -m gpu HMM tests, timing of one long string (default and MAUVE_CUDA_HMM_FP64=1), optional A/B of bk_group3 variants, one ncu capture of the chain kernel
# variants of bk_group3 on the headline workload
set -u
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
O=gpurun_out/${TAG:-hmm}
mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests -m gpu -q -x -p no:cacheprovider --timeout 300 -k "hmm" -s > $O/pytest_hmm.log 2>&1
echo "rc=$?" >> $O/pytest_hmm.log
timeout 300 python tools/hmm_time.py > $O/hmm_time.log 2>&1
MAUVE_CUDA_HMM_FP64=1 timeout 300 python tools/hmm_time.py > $O/hmm_time_fp64.log 2>&1
SHORT="--steps 10 --warmup 3 --no-dp --no-cpu --no-buildindex --no-sml"
for v in ${VARIANTS:-}; do
  MAUVE_CUDA_GROUP_VARIANT=$v timeout 300 python bench.py $SHORT > $O/bench_var$v.json 2> $O/bench_var$v.err
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'hmm_exact_chain_warp_kernel' -c 1 -o $O/hmm_warp -f python tools/hmm_time.py --one 500000 > /dev/null 2> $O/hmm_ncu.err
echo done
