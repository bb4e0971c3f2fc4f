#!/bin/bash
# Round-2 GPU session 2 (TWO B200s of one box): the sharded step on real ranks -- parity against the single-GPU rows and the
# reference's golden list first, then the bench line at N = 2 the way the driver launches it.  Everything under its own timeout.
set -u
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
O=gpurun_out/s2
mkdir -p $O
export PYTHONUNBUFFERED=1
nvidia-smi -L > $O/gpus.txt 2>&1
nvidia-smi topo -m > $O/topo.txt 2>&1
export MAUVE_CUDA_GROUP_V1=1
timeout 300 python tools/multi_gpu_check.py --gpus 2 --mbp 5 --steps 5 > $O/check_5mbp.json 2> $O/check_5mbp.err
echo "rc=$?" >> $O/check_5mbp.err
timeout 400 python tools/multi_gpu_check.py --gpus 2 --mbp 100 --steps 8 --port 29593 > $O/check_100mbp.json 2> $O/check_100mbp.err
echo "rc=$?" >> $O/check_100mbp.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu --no-buildindex --no-sml > $O/bench_n2.json 2> $O/bench_n2.err
echo "rc=$?" >> $O/bench_n2.err
echo done
