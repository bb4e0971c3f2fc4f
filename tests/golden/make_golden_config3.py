"""Mints tests/golden/config3_rows.json: the match list of the FULL BASELINE config 3 pair (synthetic 100 Mbp pair, default
weight 19) computed by the REFERENCE's own code (oracle/_ref/libmauve_ref.so = unmodified /root/reference sources, recipe
oracle/Makefile.ref): row count, sha1 of the int64 [n, 3] rows in GetMatchList order, reverse-strand rows, sum of lengths,
MemCollisionCount, and the rows the C restatement (oracle/libmauve_oracle.so) leaves out or adds (the MER_REPEAT_LIMIT
divergence of DESIGN.md section 2: expected none on this pair, which has no mer with more than 1000 copies).

Run in the build container only (needs /root/reference; ~5 minutes and ~10 GB of host memory per checker):
    python tests/golden/make_golden_config3.py
The GPU test (tests/test_zz_fullsize_gpu.py) and bench.py assert the sha1 on every run of the headline configuration.
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _oracle  # noqa: E402
from mauve_py_b200 import synth  # noqa: E402


def describe(rows):
    rows = np.ascontiguousarray(rows, dtype=np.int64)
    return {"rows": int(rows.shape[0]), "sha1": hashlib.sha1(rows.tobytes()).hexdigest(), "reverse_rows": int((rows[:, 2] < 0).sum()),
            "sum_len": int(rows[:, 0].sum()), "max_len": int(rows[:, 0].max()), "first": rows[:3].tolist(), "last": rows[-3:].tolist()}


def main():
    a, b = synth.config3_pair()
    ab, bb = a.tobytes(), b.tobytes()
    ref = _oracle.ref_checker()
    weight = ref.default_seed_weight((len(ab) + len(bb)) // 2)
    seed = ref.get_seed(weight, 3)
    out = {"workload": "synth.config3_pair() (seed 20261018, 100,000,000 bp ancestor)", "n0": len(ab), "n1": len(bb), "seed_weight": weight,
           "seed_pattern": hex(seed)}
    t0 = time.perf_counter()
    rrows, rstats = ref.find_mums(ab, bb, seed, 0)
    out["reference_seconds"] = round(time.perf_counter() - t0, 1)
    out["reference"] = describe(rrows)
    out["reference"]["collisions"] = int(rstats[0])
    print(json.dumps(out["reference"]), flush=True)
    if "--no-oracle" not in sys.argv:
        orc = _oracle.oracle_checker()
        t0 = time.perf_counter()
        orows, ostats = orc.find_mums(ab, bb, seed, 0)
        out["oracle_seconds"] = round(time.perf_counter() - t0, 1)
        out["oracle"] = describe(orows)
        out["oracle"]["repeat_limit_flag"] = int(ostats[3]) if len(ostats) > 3 else None
        same = rrows.shape == orows.shape and bool(np.array_equal(rrows, orows))
        out["oracle_equals_reference"] = same
        if not same:
            rs = {tuple(r) for r in rrows.tolist()}
            os_ = {tuple(r) for r in orows.tolist()}
            out["rows_only_in_reference"] = sorted(rs - os_)[:50]
            out["rows_only_in_oracle"] = sorted(os_ - rs)[:50]
            out["n_only_in_reference"] = len(rs - os_)
            out["n_only_in_oracle"] = len(os_ - rs)
    with open(os.path.join(HERE, "config3_rows.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps({k: v for k, v in out.items() if k not in ("reference", "oracle")}))


if __name__ == "__main__":
    main()
