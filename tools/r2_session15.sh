#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
O=gpurun_out/${TAG:-s15}
mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider --timeout 300 -k "sharded or sml" > $O/pytest_quick.log 2>&1
echo "rc=$?" >> $O/pytest_quick.log
( time timeout 1200 python bench.py ) > $O/bench_default.json 2> $O/bench_default.err
echo "rc=$?" >> $O/bench_default.err
timeout 120 python __graft_entry__.py smoke > $O/smoke.log 2>&1
echo "rc=$?" >> $O/smoke.log
echo done
