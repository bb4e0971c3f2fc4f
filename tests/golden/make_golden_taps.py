"""Mints tests/golden/gaps_mds42_calls.npz and hmm_mds42_call.npz: the gap searches and the HMM call of the REAL pipeline.

oracle/_ref/progressiveMauve_trace is the unmodified reference binary with link-time taps (oracle/trace_taps.cpp).  Aligning the
MDS42 pair with it records every MemHash::FindMatches call of recursive anchoring (pairwiseAnchorSearch, LM/ProgressiveAligner.cpp
:590-679: the two gap sequences, the seed pattern and the matches MemHash returned) and the homology HMM call of the backbone
stage (LM/Islands.h:161: the column string over '1'..'8', the 21 parameters, the H/N prediction).

    python tests/golden/make_golden_taps.py
"""
import gzip
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
BINARY = os.path.join(ROOT, "oracle", "_ref", "progressiveMauve_trace")
sys.path.insert(0, os.path.join(ROOT, "tests"))

HMM_PREFIX = 1_000_000


def main():
    import _oracle
    ref = _oracle.ref_checker()
    work = tempfile.mkdtemp()
    try:
        for name in ("mds42_recoded", "mds42_full"):
            with gzip.open(os.path.join(HERE, name + ".fa.gz"), "rb") as f, open(os.path.join(work, name + ".fa"), "wb") as g:
                shutil.copyfileobj(f, g)
        env = dict(os.environ, MAUVE_MH_TRACE=os.path.join(work, "mh.trace"), MAUVE_HMM_TRACE=os.path.join(work, "hmm.trace"))
        subprocess.check_call([BINARY, "--output=x.xmfa", "mds42_recoded.fa", "mds42_full.fa"], cwd=work, env=env,
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        mh = open(os.path.join(work, "mh.trace")).read().split("\n")
        hmm = [l for l in open(os.path.join(work, "hmm.trace")).read().split("\n") if l]
    finally:
        shutil.rmtree(work, ignore_errors=True)
    # ---- gap searches: class MemHash with the MUM tolerances (gap_mh), sequences logged ----
    gaps, kinds, i = [], {}, 0
    while i < len(mh):
        if not mh[i].startswith("@"):
            i += 1
            continue
        _, cls, rt, et, nseq, seed, l0, l1, nm = mh[i].split(" ")
        kinds[cls] = kinds.get(cls, 0) + 1
        i += 1
        if int(nseq) == 2 and int(l0) <= 2000000 and int(l1) <= 2000000:
            rows = [tuple(int(x) for x in mh[i + 2 + k].split()) for k in range(int(nm))]
            if cls.endswith("7MemHashE") and (int(rt), int(et)) == (0, 1):
                gaps.append((mh[i], mh[i + 1], int(seed), rows))
            i += 2 + int(nm)
    s0 = "\n".join(g[0] for g in gaps)
    s1 = "\n".join(g[1] for g in gaps)
    seeds = np.array([g[2] for g in gaps], dtype=np.uint64)
    counts = np.array([len(g[3]) for g in gaps], dtype=np.int64)
    rows = np.array([r for g in gaps for r in g[3]], dtype=np.int64).reshape(-1, 3)
    meta = {"find_matches_calls_by_class": kinds, "gap_searches": len(gaps), "matches": int(counts.sum()),
            "gap_len0_percentiles_0_10_50_90_99_100": [int(x) for x in np.percentile([len(g[0]) for g in gaps], [0, 10, 50, 90, 99, 100])]}
    np.savez_compressed(os.path.join(HERE, "gaps_mds42_calls.npz"), seq0=np.frombuffer(s0.encode(), dtype=np.uint8),
                        seq1=np.frombuffer(s1.encode(), dtype=np.uint8), seeds=seeds, counts=counts, rows=rows, meta=np.array(repr(meta)))
    print(meta)
    # ---- HMM: one call per LCB pair; keep a prefix of the (single, genome-sized) string, predicted by the reference's run() ----
    f = hmm[0].split(" ")
    seq, pred, params = f[1], f[2], np.array([float(x) for x in f[3:24]], dtype=np.float64)
    prefix = seq[:HMM_PREFIX].encode()
    ppred, _ = ref.hmm_run(prefix, params)
    hmeta = {"run_calls": len(hmm), "columns_per_call": [int(l.split(" ", 1)[0]) for l in hmm], "homologous_columns_first_call": pred.count("H"),
             "prefix_columns": len(prefix), "prefix_homologous": ppred.count(b"H")}
    np.savez_compressed(os.path.join(HERE, "hmm_mds42_call.npz"), sym=np.frombuffer(prefix, dtype=np.uint8),
                        pred=np.frombuffer(ppred, dtype=np.uint8), params=params, meta=np.array(repr(hmeta)))
    print(hmeta)


if __name__ == "__main__":
    main()
