#!/bin/bash
# Round-2 GPU session (EIGHT B200s of one box): the sharded step at N = 8 -- parity, per-rank stage times, the bench line the way the
# driver launches it (DP / HMM objects included)
set -u
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
O=gpurun_out/${TAG:-n8}
mkdir -p $O
export PYTHONUNBUFFERED=1
nvidia-smi -L > $O/gpus.txt 2>&1
timeout 240 python tools/multi_gpu_check.py --gpus 8 --mbp 100 --steps 8 --port 29811 > $O/check_n8.json 2> $O/check_n8.err
echo "rc=$?" >> $O/check_n8.err
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29823 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu --no-buildindex --no-sml > $O/bench_n8.json 2> $O/bench_n8.err
echo "rc=$?" >> $O/bench_n8.err
echo done
