#!/bin/bash
# A/B of environment switches on the headline step: every line of $AB (semicolon separated, e.g. "base;MAUVE_CUDA_S2_MINB=4") is one bench run
set -u
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
O=gpurun_out/${TAG:-ab}
mkdir -p $O
export PYTHONUNBUFFERED=1
SHORT="--steps 10 --warmup 3 --no-dp --no-cpu --no-buildindex --no-sml"
IFS=';' read -ra RUNS <<< "${AB:-base}"
i=0
for r in "${RUNS[@]}"; do
  i=$((i+1))
  if [ "$r" = "base" ]; then timeout 300 python bench.py $SHORT > $O/bench_$i.json 2> $O/bench_$i.err
  else env $r timeout 300 python bench.py $SHORT > $O/bench_$i.json 2> $O/bench_$i.err; fi
  echo "$r" > $O/bench_$i.name
done
if [ -n "${PYTEST_K:-}" ]; then
  timeout 600 python -m pytest tests -m gpu -q -x -p no:cacheprovider --timeout 300 -k "$PYTEST_K" > $O/pytest.log 2>&1
  echo "rc=$?" >> $O/pytest.log
fi
echo done
