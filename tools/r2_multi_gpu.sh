#!/bin/bash
# Round-2 GPU session 8 (FOUR B200s of one box): the sharded step on real ranks at N = 2 and 4 -- parity against the single-GPU rows and
# the reference's golden list, per-rank stage times, then the bench line the way the driver launches it
set -u
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
O=gpurun_out/${TAG:-s8}
mkdir -p $O
export PYTHONUNBUFFERED=1
nvidia-smi -L > $O/gpus.txt 2>&1
P=29700
for n in ${NS:-2 4}; do
  P=$((P+7))
  timeout 300 python tools/multi_gpu_check.py --gpus $n --mbp 100 --steps 8 --port $P > $O/check_n$n.json 2> $O/check_n$n.err
  echo "rc=$?" >> $O/check_n$n.err
  P=$((P+7))
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $P bench.py --gpus $n --steps 10 --warmup 3 --no-cpu --no-buildindex --no-sml --no-dp > $O/bench_n$n.json 2> $O/bench_n$n.err
  echo "rc=$?" >> $O/bench_n$n.err
done
echo done
