"""Mints the golden vectors under tests/golden/ by running the REFERENCE's own code
(oracle/_ref/libmauve_ref.so = unmodified /root/reference sources, recipe oracle/Makefile.ref).

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py
Outputs (committed):
  seeds.json            getSeed / getSeedLength / getSeedWeight / getDefaultSeedWeight tables
  sml_small.npz         DNAMemorySML::Create + Read on small synthetic genomes (inputs stored too)
  mums_small.npz        PairwiseMatchFinder / MemHash FindMatches on small synthetic pairs
  mds42_*.fa.gz         the reference's own fixture genomes (tests/mds42_*.fa, sequence lines only)
  mums_mds42.npz        the 29,403-row match list of BASELINE config 1 + counts for three rank-0 seeds
  nw_small.npz          ProfileProfile -> NWSmall -> BitTraceBack paths for random DNA pairs
  hmm_small.npz         run() predictions + Forward*Backward/P posteriors
"""
import gzip
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _oracle  # noqa: E402
from mauve_py_b200 import synth  # noqa: E402

REFDATA = "/root/reference/tests"


def canon_ties(pos, mer):
    """std::sort leaves ties in unspecified order: canonicalise positions ascending inside equal-mer runs"""
    order = np.lexsort((pos, mer))
    return pos[order]


def read_fasta(path):
    with open(path, "rb") as f:
        lines = f.read().split(b"\n")
    return b"".join(l.strip() for l in lines if l and not l.startswith(b">"))


def main():
    ref = _oracle.ref_checker()
    # ---- seeds ----
    table = {"get_seed": {}, "seed_length": {}, "seed_weight": {}, "default_weight": {}}
    for w in range(0, 34):
        for r in list(range(0, 8)) + [0x7FFFFFFF]:
            s = ref.get_seed(w, r)
            table["get_seed"]["%d,%d" % (w, r)] = s
            table["seed_length"][str(s)] = ref.seed_length(s)
            table["seed_weight"][str(s)] = ref.seed_weight(s)
    for n in [0, 1, 5, 31, 32, 33, 100, 181, 182, 1000, 4095, 4096, 4097, 10**4, 10**5, 10**6, 3976195, 3981477, 5 * 10**6, 10**8, 10**9,
              2**32 - 1]:
        table["default_weight"][str(n)] = ref.default_seed_weight(n)
    with open(os.path.join(HERE, "seeds.json"), "w") as f:
        json.dump(table, f, indent=0, sort_keys=True)

    # ---- SML ----
    out = {}
    cases = []
    rng = synth.rng_for(11)
    seqs = {
        "rand6k": synth.random_genome(6000, 0.5, rng).tobytes(),
        "lowcomplex": (b"ACGT" * 300 + b"A" * 500 + b"GATTACA" * 200 + synth.random_genome(800, 0.3, rng).tobytes()),
        "iupac": b"ACGTNNNNRYKMSWBDHVacgtnACGTTTGACCAGTNNACGATCGATCGACTAGCTAGCTAGCATCGATCGATCAGCTAGCTAGCTAGCATCG" * 6,
        "tiny": b"ACGTACGTACGTAGCTAGCTAGCATCGA",
    }
    for name, s in seqs.items():
        out["seq_" + name] = np.frombuffer(s, dtype=np.uint8)
    for name, w, r in [("rand6k", 11, 0), ("rand6k", 15, 3), ("rand6k", 19, 0), ("rand6k", 21, 2), ("rand6k", 5, 0), ("rand6k", 31, 0),
                       ("lowcomplex", 9, 0), ("lowcomplex", 15, 3), ("lowcomplex", 16, 1), ("iupac", 7, 0), ("iupac", 13, 3), ("tiny", 5, 0),
                       ("tiny", 9, 0)]:
        seed = ref.get_seed(w, r)
        pos, mer = ref.sml_build(seqs[name], seed)
        key = "%s_w%d_r%d" % (name, w, r)
        out["pos_" + key] = canon_ties(pos, mer)
        out["mer_" + key] = mer
        cases.append([name, w, r, int(seed)])
    out["cases"] = np.array(json.dumps(cases))
    np.savez_compressed(os.path.join(HERE, "sml_small.npz"), **out)

    # ---- MUMs (small) ----
    out, cases = {}, []
    for i, (n, w, r, rule, kw) in enumerate([
            (30000, 11, 0, 0, {}), (30000, 11, 0, 1, {}), (30000, 15, 3, 0, {}), (50000, 13, 0, 0, dict(snp=0.05)),
            (50000, 9, 0, 0, dict(snp=0.01, n_inv=3)), (20000, 19, 0, 0, dict(snp=0.002)), (8000, 7, 0, 1, dict(snp=0.03)),
            (40000, 21, 0, 0, dict(snp=0.001, n_inv=2)), (3000, 5, 0, 0, dict(snp=0.02))]):
        a, b = synth.small_pair(n, seed=100 + i, **kw)
        seed = ref.get_seed(w, r)
        rows, stats = ref.find_mums(a, b, seed, rule)
        out["a%d" % i] = np.frombuffer(a, dtype=np.uint8)
        out["b%d" % i] = np.frombuffer(b, dtype=np.uint8)
        out["rows%d" % i] = rows
        cases.append([i, w, r, rule, int(seed), int(stats[0]), int(stats[1])])
        print("mums case", i, "n", n, "w", w, "->", rows.shape[0], "matches", "collisions", int(stats[0]))
    # degenerate inputs: palindromic / repeats / short
    extra = [(b"ACGT" * 50, b"ACGT" * 60, 5, 0), (b"A" * 300, b"A" * 200, 5, 0), (b"ACGTTGCAAGCT", b"ACGTTGCAAGCT", 5, 0),
             (b"ACG", b"ACGTACGT", 5, 0), (b"", b"ACGTACGTAA", 5, 0)]
    for j, (a, b, w, r) in enumerate(extra):
        i = 100 + j
        seed = ref.get_seed(w, r)
        rows, stats = ref.find_mums(a, b, seed, 0)
        out["a%d" % i] = np.frombuffer(a, dtype=np.uint8)
        out["b%d" % i] = np.frombuffer(b, dtype=np.uint8)
        out["rows%d" % i] = rows
        cases.append([i, w, r, 0, int(seed), int(stats[0]), int(stats[1])])
    out["cases"] = np.array(json.dumps(cases))
    np.savez_compressed(os.path.join(HERE, "mums_small.npz"), **out)

    # ---- MDS42 (BASELINE config 1) ----
    g0 = read_fasta(os.path.join(REFDATA, "mds42_recoded.fa"))
    g1 = read_fasta(os.path.join(REFDATA, "mds42_full.fa"))
    for name, g in (("mds42_recoded", g0), ("mds42_full", g1)):
        with gzip.GzipFile(os.path.join(HERE, name + ".fa.gz"), "wb", compresslevel=9, mtime=0) as f:
            f.write(b">" + name.encode() + b"\n" + g + b"\n")
    out = {}
    seed = ref.get_seed(15, 3)
    rows, stats = ref.find_mums(g0, g1, seed, 0)
    txt = "".join("%d\t%d\t%d\n" % (r[0], r[1], r[2]) for r in rows).encode()
    out["rows_w15_r3"] = rows
    meta = {"w15_r3": {"seed": int(seed), "n": int(rows.shape[0]), "sum_len": int(rows[:, 0].sum()), "reverse": int((rows[:, 2] < 0).sum()),
                       "collisions": int(stats[0]), "md5_len_start0_start1": hashlib.md5(txt).hexdigest()}}
    print("mds42 w15 r3:", meta["w15_r3"])
    for w, r in [(15, 0), (11, 0), (21, 0)]:
        seed = ref.get_seed(w, r)
        rows, stats = ref.find_mums(g0, g1, seed, 0)
        meta["w%d_r%d" % (w, r)] = {"seed": int(seed), "n": int(rows.shape[0]), "sum_len": int(rows[:, 0].sum()),
                                    "sha1_rows": hashlib.sha1(rows.tobytes()).hexdigest()}
        print("mds42", w, r, meta["w%d_r%d" % (w, r)])
    # SML digest of one genome (positions canonicalised)
    pos, mer = ref.sml_build(g1, ref.get_seed(15, 3))
    meta["sml_full_w15_r3"] = {"n": int(mer.size), "sha1_mer": hashlib.sha1(mer.tobytes()).hexdigest(),
                               "sha1_pos_canon": hashlib.sha1(canon_ties(pos, mer).tobytes()).hexdigest()}
    out["meta"] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(HERE, "mums_mds42.npz"), **out)

    # ---- NW ----
    rng = synth.rng_for(77)
    pairs = synth.dp_pairs(150, 1, 700, seed=78)
    pairs += [(b"A", b"A"), (b"A", b"C"), (b"A", b"ACGTACGT"), (b"ACGTACGT", b"T"), (b"AC", b"CA"), (b"ACGT" * 70, b"ACGT" * 64),
              (b"A" * 300, b"A" * 257), (b"ACGTTGCATGCATGCAAGT" * 14, b"TTTTTTTTTT"), (b"G" * 256, b"G" * 256), (b"C" * 257, b"C" * 255)]
    for _ in range(20):  # unrelated pairs: long gaps, many ties
        la, lb = int(rng.integers(1, 400)), int(rng.integers(1, 400))
        pairs.append((synth.random_genome(la, 0.5, rng).tobytes(), synth.random_genome(lb, 0.5, rng).tobytes()))
    out = {"n": np.array(len(pairs))}
    for i, (a, b) in enumerate(pairs):
        path, _ = ref.nw_align(a, b)
        out["a%d" % i] = np.frombuffer(a, dtype=np.uint8)
        out["b%d" % i] = np.frombuffer(b, dtype=np.uint8)
        out["p%d" % i] = np.frombuffer(path, dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "nw_small.npz"), **out)

    # ---- HMM ----
    out, cases = {}, []
    for i, (n, gc, goh, gou, pid, sd) in enumerate([(3000, 0.5, 1e-5, 1e-9, 0.7, 1), (10000, 0.41, 0.0, 0.0, 0.0, 2), (1, 0.5, 1e-5, 1e-9, 0.7, 3),
                                                    (65, 0.6, 1e-5, 1e-9, 0.7, 4), (20000, 0.508, 1e-5, 1e-9, 0.7, 5), (64, 0.5, 1e-4, 1e-8, 0.9, 6)]):
        sym = synth.hmm_string(n, seed=sd, block=300)
        params = ref.hmm_params(gc, goh, gou, pid)
        pred, post = ref.hmm_run(sym, params)
        out["sym%d" % i] = np.frombuffer(sym, dtype=np.uint8)
        out["params%d" % i] = params
        out["pred%d" % i] = np.frombuffer(pred, dtype=np.uint8)
        out["post%d" % i] = post
        cases.append([i, gc, goh, gou, pid])
    out["cases"] = np.array(json.dumps(cases))
    np.savez_compressed(os.path.join(HERE, "hmm_small.npz"), **out)
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
