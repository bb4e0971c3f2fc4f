"""Mints tests/golden/dp_mds42_calls.npz: gapped-DP inputs and paths from the REAL pipeline.

oracle/_ref/progressiveMauve_trace is the unmodified reference binary with a link-time tap on muscle::GlobalAlign
(oracle/trace_taps.cpp).  Aligning the MDS42 pair with it records every GlobalAlign call: the two profiles (as letter
strings when every column is one ungapped ACGT letter) and the path NWSmall + BitTraceBack returned.  The fixture keeps the 200
largest calls and 1,300 random ones, plus the statistics of the whole run (they describe the DP workload buildIndex really
generates: DESIGN.md section 5).

    python tests/golden/make_golden_dp.py
"""
import gzip
import os
import shutil
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
BINARY = os.path.join(ROOT, "oracle", "_ref", "progressiveMauve_trace")


def main():
    work = tempfile.mkdtemp()
    try:
        for name in ("mds42_recoded", "mds42_full"):
            with gzip.open(os.path.join(HERE, name + ".fa.gz"), "rb") as f, open(os.path.join(work, name + ".fa"), "wb") as g:
                shutil.copyfileobj(f, g)
        env = dict(os.environ, MAUVE_DP_TRACE=os.path.join(work, "dp.trace"))
        subprocess.check_call([BINARY, "--output=x.xmfa", "mds42_recoded.fa", "mds42_full.fa"], cwd=work, env=env,
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        calls = [l.rstrip("\n").split(" ") for l in open(os.path.join(work, "dp.trace"))]
    finally:
        shutil.rmtree(work, ignore_errors=True)
    ok = [c for c in calls if c[0] != "-"]
    la = np.array([len(c[0]) for c in ok], dtype=np.int64)
    lb = np.array([len(c[1]) for c in ok], dtype=np.int64)
    cells = la * lb
    meta = {"calls": len(calls), "single_sequence_acgt_calls": len(ok), "cells": int(cells.sum()), "max_cells": int(cells.max()),
            "lenA_percentiles_0_10_50_90_99_100": [int(x) for x in np.percentile(la, [0, 10, 50, 90, 99, 100])]}
    order = np.argsort(-cells, kind="stable")
    pick = set(order[:200].tolist())
    rng = np.random.default_rng(20261017)
    pick |= set(rng.choice(len(ok), 1300, replace=False).tolist())
    pick = sorted(pick)
    a = "\n".join(ok[i][0] for i in pick)
    b = "\n".join(ok[i][1] for i in pick)
    p = "\n".join(ok[i][2] for i in pick)
    np.savez_compressed(os.path.join(HERE, "dp_mds42_calls.npz"), a=np.frombuffer(a.encode(), dtype=np.uint8),
                        b=np.frombuffer(b.encode(), dtype=np.uint8), path=np.frombuffer(p.encode(), dtype=np.uint8), meta=np.array(repr(meta)))
    print(meta, len(pick))


if __name__ == "__main__":
    main()
