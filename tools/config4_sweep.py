#!/usr/bin/env python
"""BASELINE config 4: spaced-seed weight sweep on a large synthetic genome pair -- SML build and match ENUMERATION throughput.

    python tools/config4_sweep.py [--gbp 1.0] [--weights 11-21] [--ranks 0,3] [--reps 2]

SURVEY.md 8d "C4": seed 20261019; ancestor of --gbp * 1e9 bases + a 1 % SNP copy; weights 11..21 at seed rank 0 and 3; stages: sorted mer
list build (mcu_sml_build with no host outputs: H2D of the genome + pack + seed generation + radix sort) and enumeration of the unique
seed pairs (mcu_session_enumerate on resident genomes; no extension).  One JSON line per (weight, rank); run on the GPU box
(tools/gpu_session.sh does not include it: ~1-2 minutes at 1 Gbp, 60 GB of HBM).  Needs ~6 x --gbp GB of host memory for the genomes.
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gbp", type=float, default=1.0)
    ap.add_argument("--weights", default="11-21")
    ap.add_argument("--ranks", default="0,3")
    ap.add_argument("--reps", type=int, default=2)
    args = ap.parse_args()
    import mauve_py_b200 as mp
    from mauve_py_b200 import synth
    from mauve_py_b200._capi import check
    lib = mp.lib()
    check(lib.mcu_init(0))
    n = int(args.gbp * 1e9)
    rng = synth.rng_for(20261019)
    step = 50_000_000   # generated in slices: the float32 draw of random_genome is 4 bytes per base
    a = np.concatenate([synth.random_genome(min(step, n - o), 0.5, rng) for o in range(0, n, step)])
    b = np.concatenate([synth.snps(a[o:o + step], 0.01, rng) for o in range(0, n, step)])
    lo, hi = (int(x) for x in args.weights.split("-"))
    sess = mp.AnchorSession()
    sess.upload(a, b)
    for w in range(lo, hi + 1):
        for rank in (int(x) for x in args.ranks.split(",")):
            seed = mp.getSeed(w, rank)
            L, wt = mp.getSeedLength(seed), mp.getSeedWeight(seed)
            sml_ms, enum_ms = [], []
            for _ in range(args.reps):
                t0 = time.perf_counter()
                for g in (a, b):
                    out_len = C.c_uint64(0)
                    check(lib.mcu_sml_build(g.ctypes.data, g.size, seed, None, None, None, C.byref(out_len)))
                sml_ms.append(1e3 * (time.perf_counter() - t0))
                t0 = time.perf_counter()
                sess.enumerate(seed)
                enum_ms.append(1e3 * (time.perf_counter() - t0))
            nbases = 2 * n
            print(json.dumps({"config": "C4", "genome_bp": n, "weight_requested": w, "rank": rank, "seed": hex(seed), "seed_length": L, "seed_weight": wt,
                              "sml_build_ms_both_genomes": min(sml_ms), "sml_build_mbp_s": nbases / 1e6 / (min(sml_ms) * 1e-3),
                              "enumerate_ms": min(enum_ms), "enumerate_mbp_s": nbases / 1e6 / (min(enum_ms) * 1e-3),
                              "note": "sml build includes the H2D copy of the genome (host buffers); enumeration runs on resident genomes"}), flush=True)


if __name__ == "__main__":
    main()
