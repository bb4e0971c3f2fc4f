"""Mints tests/golden/sol_small.npz and sol_mds42.npz by running the REFERENCE's own SeedOccurrenceList::construct and
GetPairwiseAnchorScore (oracle/_ref/libmauve_ref_full.so = unmodified /root/reference sources, recipe oracle/Makefile.ref).

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_sol.py
  sol_small.npz   small synthetic sequences / pairs (inputs stored): frequencies per seed, anchor scores per LCB for the
                  reference's match list cut into arbitrary LCBs, with and without penalize_repeats
  sol_mds42.npz   BASELINE config 1 (the MDS42 pair, coding seed w15): sha1 of the frequency arrays + every 199th value + the values around the maximum, and the anchor scores of the 29,403-row golden match list in LCBs of 64 rows
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _golden  # noqa: E402
import _oracle  # noqa: E402
from mauve_py_b200 import synth  # noqa: E402


def iupac(seq, k, rng):
    s = bytearray(seq)
    for i in rng.integers(0, len(s), k):
        s[i] = int(rng.choice(list(b"NRYKMSWBVDHnacgtXx")))
    return bytes(s)


def small_cases():
    rng = np.random.default_rng(20261021)
    a, b = synth.small_pair(40000, seed=7)
    unit = synth.random_genome(400, 0.5, rng).tobytes()
    core = synth.random_genome(6000, 0.45, rng).tobytes()
    inv = synth.revcomp(np.frombuffer(core[2000:4000], dtype=np.uint8)).tobytes()
    x = unit + core[:3000] + unit + core[3000:] + unit + b"A" * 300 + b"ACAC" * 100
    y = core[:2000] + inv + unit + core[4000:] + unit + b"A" * 200
    return {"snp_pair": (a, b), "iupac_pair": (iupac(a, 300, rng), iupac(b, 300, rng)), "repeats_inversion": (x, y),
            "short": (b"ACGTTGCAACGTACGTTTGACCA" * 3, b"ACGTTGCAACGTACGTTTGACCA" * 2 + b"GATTACA")}


def main():
    ref = _oracle.ref_checker()
    rng = np.random.default_rng(5)
    out, cases = {}, []
    for name, (s0, s1) in small_cases().items():
        for w, rank in ((11, 0), (15, 3), (9, 0), (19, 3)):
            seed = ref.get_seed(w, rank)
            if min(len(s0), len(s1)) < ref.seed_length(seed):
                continue
            key = "%s_w%d_r%d" % (name, w, rank)
            rows, _ = ref.find_mums(s0, s1, seed, 0)
            n = rows.shape[0]
            cuts = np.unique(np.concatenate([[0, n], rng.integers(0, n + 1, 6)])).astype(np.uint64)
            out[key + "_f0"] = _oracle.sol_build(s0, seed, use_ref=True)
            out[key + "_f1"] = _oracle.sol_build(s1, seed, use_ref=True)
            out[key + "_rows"] = rows
            out[key + "_cuts"] = cuts
            out[key + "_lcb"] = _oracle.anchor_scores(s0, s1, seed, rows, cuts, False, use_ref=True)[0]
            out[key + "_lcb_pen"] = _oracle.anchor_scores(s0, s1, seed, rows, cuts, True, use_ref=True)[0]
            cases.append({"key": key, "name": name, "w": w, "rank": rank, "seed": seed})
    for name, (s0, s1) in small_cases().items():
        out["seq_%s_0" % name] = np.frombuffer(s0, dtype=np.uint8)
        out["seq_%s_1" % name] = np.frombuffer(s1, dtype=np.uint8)
    out["cases"] = json.dumps(cases)
    np.savez_compressed(os.path.join(HERE, "sol_small.npz"), **out)
    print("sol_small.npz: %d cases" % len(cases))

    # ---- MDS42 (BASELINE config 1) ----
    g0, g1 = _golden.mds42()
    z = _golden.npz("mums_mds42.npz")
    rows = z["rows_w15_r3"]
    seed = ref.get_seed(15, 3)
    f0 = _oracle.sol_build(g0, seed, use_ref=True)
    f1 = _oracle.sol_build(g1, seed, use_ref=True)
    cuts = np.arange(0, rows.shape[0] + 64, 64, dtype=np.uint64)
    cuts[-1] = rows.shape[0]
    lcb = _oracle.anchor_scores(g0, g1, seed, rows, cuts, False, use_ref=True)[0]
    md = {"seed": seed, "sha1_f0": hashlib.sha1(f0.tobytes()).hexdigest(), "sha1_f1": hashlib.sha1(f1.tobytes()).hexdigest(),
          "n0": len(g0), "n1": len(g1), "lcb_rows": 64}
    nz0, nz1 = np.flatnonzero(f0 != 1.0), np.flatnonzero(f1 != 1.0)
    md["not_one_0"], md["not_one_1"] = int(nz0.size), int(nz1.size)
    md["sum_0"], md["sum_1"] = float(f0.astype(np.float64).sum()), float(f1.astype(np.float64).sum())
    # the arrays themselves are 16 MB each: the fixture keeps their sha1, every 199th value and the values around the maximum
    i0 = np.unique(np.concatenate([np.arange(0, f0.size, 199), np.arange(max(int(f0.argmax()) - 200, 0), min(int(f0.argmax()) + 200, f0.size))]))
    i1 = np.unique(np.concatenate([np.arange(0, f1.size, 199), np.arange(max(int(f1.argmax()) - 200, 0), min(int(f1.argmax()) + 200, f1.size))]))
    np.savez_compressed(os.path.join(HERE, "sol_mds42.npz"), idx0=i0.astype(np.uint32), val0=f0[i0], idx1=i1.astype(np.uint32), val1=f1[i1],
                        cuts=cuts, lcb=lcb, meta=json.dumps(md))
    print("sol_mds42.npz: %d / %d positions differ from 1, %d LCBs, total %.0f" % (nz0.size, nz1.size, lcb.size, lcb.sum()))


if __name__ == "__main__":
    main()
