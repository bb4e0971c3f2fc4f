#!/usr/bin/env python
"""BASELINE config 5: gapped-DP stress -- many inter-anchor region pairs 100 bp - 10 kbp, plus one HomologyHMM string per region.

    python tools/config5_dp.py [--regions 100000] [--chunk 20000]

SURVEY.md 8d "C5": seed 20261020; lenA log-uniform in [100, 10000], B = A with 5 % SNPs and 1 % indel events.  The full configuration is
1,000,000 regions (~1.1e13 cells: about half a minute of device time at the measured 372 GCUPS; generating the regions in Python takes
longer than aligning them, hence the default of 100,000).  Regions go to the device in chunks of --chunk; GCUPS counts full-matrix
cells over the summed device time.  The kernel is unbanded (exact by construction), so there is no band to sweep.
Prints one JSON line.
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--regions", type=int, default=100_000)
    ap.add_argument("--chunk", type=int, default=20_000)
    args = ap.parse_args()
    import mauve_py_b200 as mp
    from mauve_py_b200 import synth
    from mauve_py_b200._capi import check
    check(mp.lib().mcu_init(0))
    params = mp.libmems.hmm_params(0.5, 1e-5, 1e-9, 0.7)
    cells = dev_ms = hmm_cols = hmm_ms = 0.0
    t_gen = t_wall = 0.0
    done = 0
    while done < args.regions:
        k = min(args.chunk, args.regions - done)
        t0 = time.perf_counter()
        pairs = synth.dp_pairs(k, 100, 10000, seed=20261020 + done)
        sym = [synth.hmm_string(len(p[0]), seed=done + i, block=300) for i, p in enumerate(pairs)]
        t_gen += time.perf_counter() - t0
        t0 = time.perf_counter()
        res = mp.libmems.nw_batch_arrays(*synth.dp_arrays(pairs))
        _p, _q, ms = mp.run_batch(sym, params, True)
        t_wall += time.perf_counter() - t0
        cells += float(res["stats"][0])
        dev_ms += float(res["device_ms"])
        hmm_cols += float(sum(len(s) for s in sym))
        hmm_ms += float(ms)
        done += k
    print(json.dumps({"config": "C5", "regions": args.regions, "cells": cells, "dp_device_ms": dev_ms, "gcups": cells / (dev_ms * 1e-3) / 1e9,
                      "hmm_columns": hmm_cols, "hmm_device_ms": hmm_ms, "hmm_columns_s": hmm_cols / (hmm_ms * 1e-3),
                      "host_wall_s_dp_plus_hmm": t_wall, "generation_s": t_gen}))


if __name__ == "__main__":
    main()
