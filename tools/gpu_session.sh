#!/bin/bash
# One gpurun call that collects everything a round needs from the GPU box (about 20-25 minutes on one B200):
#
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/gpu_session.sh'
#
# Everything lands under gpurun_out/ (merged back by gpurun); summarise what is to be judged into profiles/.
#   pytest_gpu.log        python -m pytest tests -m gpu (no -x: one failure must not hide the rest), with durations
#   bench_n1.json/.err    the bench line (N = 1) incl. dp / hmm / buildindex objects
#   launches.csv          ncu launch list of a short bench step (per-launch device time: compare SHARES)
#   bk_group.ncu-rep      ncu --set full of the dominant kernel (bk_group_kernel), + scatter passes
#   sol.ncu-rep           ncu --set full of the seed occurrence list / anchor score kernels (8f-2)
#   seams.log             reference binary vs the seam binaries on the MDS42 pair (wall seconds, XMFA sha1, seam reports)
#   dropin*.log           the C++ drop-in checks with their own timings
#   config5.json, config4.jsonl   BASELINE configs 5 (50,000 regions) and 4 (1 Gbp weight sweep)
#   ab/, ab.log           (AB=1) tools/gpu_ab_patch.sh: every experiments/*.patch against this tree
set -u
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1

timeout 1200 python -m pytest tests -m gpu -q --durations=30 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log

timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err

SHORT="python bench.py --steps 2 --warmup 1 --no-dp --no-cpu --no-buildindex"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $SHORT > /dev/null 2> gpurun_out/launches.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'bk_group_kernel|bkf_scatter1_kernel|bkf_scatter2_kernel' -s 3 -c 3 \
    -o gpurun_out/bk_group -f $SHORT > /dev/null 2> gpurun_out/bk_group.err

cat > gpurun_out/_sol_probe.py <<'PY'
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import mauve_py_b200 as mp
from mauve_py_b200 import synth
from mauve_py_b200._capi import check
check(mp.lib().mcu_init(0))
a, b = synth.config3_pair(n=20_000_000)
seed = mp.getSeed(mp.getDefaultSeedWeight(a.size), mp.CODING_SEED)
rows, _ = mp.libmems.find_mums(a.tobytes(), b.tobytes(), seed)
cuts = np.arange(0, rows.shape[0] + 64, 64, dtype=np.uint64); cuts[-1] = rows.shape[0]
for _ in range(2):
    lcb, ms = mp.libmems.anchor_scores(a.tobytes(), b.tobytes(), rows, cuts, seed=seed)
print("rows", rows.shape[0], "lcbs", lcb.size, "total", float(lcb.sum()))
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'sol_count_kernel|sol_smooth_kernel|anchor_score_kernel' -c 6 \
    -o gpurun_out/sol -f python gpurun_out/_sol_probe.py > gpurun_out/sol.log 2>&1

# the reference binary and the seam binaries on the MDS42 pair
W=$(mktemp -d)
for n in mds42_recoded mds42_full; do zcat tests/golden/$n.fa.gz > $W/$n.fa; done
{
  for bin in progressiveMauve progressiveMauve_cuda progressiveMauve_cuda_mh progressiveMauve_cuda_all; do
    for sol in 0 1; do
      [ "$sol" = 1 ] && [ "$bin" != progressiveMauve_cuda_all ] && continue
      rm -f $W/*.sslist
      s=$(date +%s.%N)
      ( cd $W && MAUVE_CUDA_GAP_SEAM=1 MAUVE_CUDA_SOL_SEAM=$sol MAUVE_CUDA_SEAM_REPORT=1 timeout 600 "$OLDPWD/oracle/_ref/$bin" --output=$bin.$sol.xmfa mds42_recoded.fa mds42_full.fa > $bin.$sol.log 2> $bin.$sol.err )
      rc=$?
      e=$(date +%s.%N)
      echo "== $bin sol_seam=$sol rc=$rc wall_s=$(awk -v a="$s" -v b="$e" 'BEGIN{printf "%.2f", b - a}') xmfa_sha1=$(grep -v '^#' $W/$bin.$sol.xmfa | sed -E 's/^(> *[^ ]+ [^ ]+) .*/\1/' | sha1sum | cut -c1-40)"
      grep -a "seam:" $W/$bin.$sol.err
    done
  done
} > gpurun_out/seams.log 2>&1
rm -rf $W

timeout 300 oracle/_ref/dropin_check gaps 5000 3 > gpurun_out/dropin_gaps.log 2>&1
timeout 300 oracle/_ref/dropin_check dp 400 3 > gpurun_out/dropin_dp.log 2>&1
timeout 300 oracle/_ref/dropin_check hmm 5000000 9 > gpurun_out/dropin_hmm.log 2>&1
# BASELINE configs 4 and 5 at (or near) full size
timeout 500 python tools/config5_dp.py --regions 50000 > gpurun_out/config5.json 2> gpurun_out/config5.err
timeout 900 python tools/config4_sweep.py --gbp 1.0 --reps 1 > gpurun_out/config4.jsonl 2> gpurun_out/config4.err
# A/B of the unverified patches under experiments/ against this tree (adds ~8 minutes): AB=1 bash tools/gpu_session.sh
[ "${AB:-0}" = 1 ] && timeout 1500 bash tools/gpu_ab_patch.sh > gpurun_out/ab.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/nvidia_smi.csv 2>&1
echo done
