#!/bin/bash
# A/B of the unverified patches under experiments/ in ONE gpurun call: the tree as it is (A) against a scratch copy with
# one patch applied (B), same box, same clocks, back to back.  Nothing in the repository is modified.
#
#   /usr/local/graft/bin/gpurun --timeout 1800 -- 'bash tools/gpu_ab_patch.sh'            # every experiments/*.patch
#   /usr/local/graft/bin/gpurun --timeout 900  -- 'bash tools/gpu_ab_patch.sh experiments/r02_dp_predicated_traceback_bits.patch'
#   AB_DRY=1 bash tools/gpu_ab_patch.sh                                                    # no GPU: copy, patch, build only
#
# Per patch, under gpurun_out/ab/<patch name>/:
#   build.log      patch + nvcc output of the scratch copy
#   tests.log      the -m gpu tests selected for the patch (AB_TESTS_<n> below), each under its own timeout
#   a.json b.json  bench lines (A = this tree, B = patched copy) from the same short bench command
#   summary.txt    ms_per_step, value, dp.gcups and the stage / kernel times of both, side by side
# A patch is kept only if tests.log is green AND b.json beats a.json on the number the patch is about.
set -u
ROOT="${GRAFT_REPO_ROOT:-$(cd "$(dirname "$0")/.." && pwd)}"
cd "$ROOT"
OUT="$ROOT/gpurun_out/ab"
mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
DRY="${AB_DRY:-0}"

# tests (-k expression over tests/test_gpu_parity.py) and bench flags per patch, chosen by a substring of the file name
tests_for() {
  case "$1" in
    *dp_*)       echo "nw";;
    *scatter1*)  echo "two_phase or sharded or mums";;
    *)           echo "not radix";;
  esac
}
bench_for() {
  case "$1" in
    *dp_*)       echo "--steps 3 --warmup 3 --no-cpu --no-buildindex";;   # keeps the dp object
    *)           echo "--steps 5 --warmup 3 --no-dp --no-cpu --no-buildindex";;
  esac
}

PATCHES=("$@")
[ ${#PATCHES[@]} -eq 0 ] && PATCHES=(experiments/*.patch)

for P in "${PATCHES[@]}"; do
  NAME=$(basename "$P" .patch)
  D="$OUT/$NAME"
  mkdir -p "$D"
  W=$(mktemp -d)
  # scratch copy: sources + the prebuilt checkers (oracle/_ref, oracle/*.so travel with the snapshot); no .git, no gpurun_out
  tar -C "$ROOT" --exclude=.git --exclude=gpurun_out --exclude=build -cf - . | tar -C "$W" -xf -
  {
    echo "== patch $P"
    ( cd "$W" && patch -p0 < "$ROOT/$P" ) || { echo "PATCH FAILED"; rm -rf "$W"; continue; }
    ( cd "$W" && python -c "from mauve_py_b200 import _build; print(_build.build_library(force=True))" )
    echo "build rc=$?"
  } > "$D/build.log" 2>&1
  grep -q "build rc=0" "$D/build.log" || { echo "$NAME: build failed (see $D/build.log)"; rm -rf "$W"; continue; }
  if [ "$DRY" = 1 ]; then echo "$NAME: patched copy builds (dry run, no GPU commands)"; rm -rf "$W"; continue; fi

  # the deadlock of the first scatter1 version is why every GPU command here sits under its own timeout
  ( cd "$W" && timeout 600 python -m pytest tests/test_gpu_parity.py -k "$(tests_for "$NAME")" -m gpu -q -x -p no:cacheprovider --timeout 120 ) > "$D/tests.log" 2>&1
  echo "pytest rc=$?" >> "$D/tests.log"
  BF=$(bench_for "$NAME")
  ( cd "$ROOT" && timeout 400 python bench.py $BF ) > "$D/a.json" 2> "$D/a.err"
  ( cd "$W" && timeout 400 python bench.py $BF ) > "$D/b.json" 2> "$D/b.err"
  case "$NAME" in *scatter1*)   # only a sharded run reaches this code: one rank's share (shard 0 of 8) of the 100 Mbp pair, kernel times per step
    ( cd "$ROOT" && timeout 300 python tools/shard_step.py 100 0 8 5 ) > "$D/a_shard0of8.txt" 2>&1
    ( cd "$W" && timeout 300 python tools/shard_step.py 100 0 8 5 ) > "$D/b_shard0of8.txt" 2>&1;;
  esac
  python - "$D" > "$D/summary.txt" 2>&1 <<'PY'
import json, sys, os
d = sys.argv[1]
def load(n):
    try:
        return json.loads(open(os.path.join(d, n)).read().strip().splitlines()[-1])
    except Exception as e:
        return {"error": repr(e)}
a, b = load("a.json"), load("b.json")
def pick(x):
    r = x.get("roofline", {}) if isinstance(x, dict) else {}
    step = r.get("step", {})
    out = {"ms_per_step": x.get("ms_per_step"), "value": x.get("value"), "e2e": (x.get("e2e") or {}).get("value"),
           "dp_gcups": (x.get("dp") or {}).get("value"), "dp_device_ms": (x.get("dp") or {}).get("device_ms"), "error": x.get("error")}
    for k, v in (step.get("kernel_ms") or {}).items():
        out["kernel." + k] = v
    return out
pa, pb = pick(a), pick(b)
print("%-28s %16s %16s" % ("", "A (tree)", "B (patched)"))
for k in sorted(set(pa) | set(pb)):
    print("%-28s %16s %16s" % (k, pa.get(k), pb.get(k)))
print(open(os.path.join(d, "tests.log")).read().strip().splitlines()[-2:])
for n in ("a_shard0of8.txt", "b_shard0of8.txt"):   # rows: matches, [pack, scatter1, scatter2, group, candidate, extend, order] ms, total
    if os.path.exists(os.path.join(d, n)):
        print(n, open(os.path.join(d, n)).read().strip().splitlines()[-1])
PY
  echo "== $NAME"; cat "$D/summary.txt"
  rm -rf "$W"
done
echo done
