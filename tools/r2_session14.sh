#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
O=gpurun_out/${TAG:-s14}
mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider --timeout 300 -k "sml" -s > $O/pytest_sml.log 2>&1
echo "rc=$?" >> $O/pytest_sml.log
for n in ${NS:-2 4}; do
  timeout 300 python tools/multi_gpu_check.py --gpus $n --mbp 100 --steps 5 --port $((29900 + n)) > $O/check_n$n.json 2> $O/check_n$n.err
  echo "rc=$?" >> $O/check_n$n.err
done
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29951 bench.py --gpus 4 --steps 10 --warmup 3 --no-cpu --no-buildindex --no-dp > $O/bench_n4.json 2> $O/bench_n4.err
echo "rc=$?" >> $O/bench_n4.err
echo done
