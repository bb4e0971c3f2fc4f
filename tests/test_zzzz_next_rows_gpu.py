"""GPU (-m gpu): the rows next to the hot path: seed occurrence list + anchor scores, through the C ABI
(mcu_sol_build, mcu_anchor_scores), against the golden vectors minted from the reference's own SeedOccurrenceList /
GetPairwiseAnchorScore, against the oracle on further inputs, through a full-size property check, and through the C++ adapters
next to the reference classes (oracle/_ref/dropin_check_next).  Written after the round's GPU budget was spent: the file sorts
last so that nothing verified earlier depends on it."""
import hashlib
import os

import numpy as np
import pytest

import _golden
import _oracle
import _properties as P
from mauve_py_b200 import synth
from _bins import NEXT_BIN, run_next, write_fasta

pytestmark = pytest.mark.gpu


def _sol(mp, seq, seed):
    sml = mp.DNAMemorySML()
    sml.Create(seq, seed)
    sol = mp.SeedOccurrenceList()
    sol.construct(sml)
    return sol


def test_sol_and_anchor_scores_golden(mp):
    z = _golden.npz("sol_small.npz")
    for c in _golden.cases(z):
        k = c["key"]
        s0, s1 = z["seq_%s_0" % c["name"]].tobytes(), z["seq_%s_1" % c["name"]].tobytes()
        sol0, sol1 = _sol(mp, s0, c["seed"]), _sol(mp, s1, c["seed"])
        assert np.array_equal(sol0.frequencies().view(np.uint32), z[k + "_f0"].view(np.uint32)), k
        assert np.array_equal(sol1.frequencies().view(np.uint32), z[k + "_f1"].view(np.uint32)), k
        assert sol0.getFrequency(0) == float(z[k + "_f0"][0])
        rows, cuts = z[k + "_rows"], z[k + "_cuts"]
        for pen, name in ((False, "_lcb"), (True, "_lcb_pen")):
            lcb, ms = mp.libmems.anchor_scores(s0, s1, rows, cuts, sol_1=sol0, sol_2=sol1, penalize_repeats=pen)
            assert np.array_equal(lcb, z[k + name]), (k, pen)
            # frequencies built inside the call from the seed
            lcb2, ms2 = mp.libmems.anchor_scores(s0, s1, rows, cuts, seed=c["seed"], penalize_repeats=pen)
            assert np.array_equal(lcb2, z[k + name]) and np.array_equal(ms, ms2), (k, pen)
            _, oms = _oracle.anchor_scores(s0, s1, c["seed"], rows, cuts, pen, freq=(z[k + "_f0"], z[k + "_f1"]))
            assert np.array_equal(ms, oms), (k, pen)


def test_sol_mds42_golden(mp):
    """BASELINE config 1: both MDS42 genomes (coding seed w15) and the anchor scores of the 29,403-row golden match list"""
    z = _golden.npz("sol_mds42.npz")
    md = _golden.meta(z)
    g0, g1 = _golden.mds42()
    sol0, sol1 = _sol(mp, g0, md["seed"]), _sol(mp, g1, md["seed"])
    assert hashlib.sha1(sol0.frequencies().tobytes()).hexdigest() == md["sha1_f0"]
    assert hashlib.sha1(sol1.frequencies().tobytes()).hexdigest() == md["sha1_f1"]
    rows = _golden.npz("mums_mds42.npz")["rows_w15_r3"]
    lcb, ms = mp.libmems.anchor_scores(g0, g1, rows, z["cuts"], sol_1=sol0, sol_2=sol1)
    assert np.array_equal(lcb, z["lcb"])
    # the reference's interface, one LCB
    first = [mp.Match(int(r[0]), [int(r[1]), int(r[2])]) for r in rows[:64]]
    assert mp.GetPairwiseAnchorScore(first, [g0, g1], None, sol0, sol1) == float(z["lcb"][0])


@pytest.mark.parametrize("w,rank", [(7, 0), (11, 0), (15, 3), (19, 3), (21, 0), (31, 0)])
def test_sol_edges_vs_oracle(mp, orc, w, rank):
    """no seed at all, exactly one seed, long runs (the bisection branch of sol_run_length), IUPAC and lower-case letters"""
    rng = np.random.default_rng(300 + w)
    seed = orc.get_seed(w, rank)
    L = orc.seed_length(seed)
    a, _b = synth.repeat_rich_pair(n=80_000, unit=61, copies=500, seed=w)
    seqs = [a, b"A" * 200_000, b"AC" * 40_000 + a[:3000], bytes(rng.choice(list(b"ACGT"), L).astype(np.uint8)),
            bytes(rng.choice(list(b"ACGT"), L + 1).astype(np.uint8)), bytes(rng.choice(list(b"ACGT"), max(L - 1, 1)).astype(np.uint8)), b"G",
            bytes(rng.choice(list(b"ACGTNRYKMSWnacgt"), 5000).astype(np.uint8))]
    for s in seqs:
        got = _sol(mp, s, seed).frequencies()
        want = _oracle.sol_build(s, seed)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (len(s), w, np.flatnonzero(got != want)[:5])


def test_anchor_scores_argument_errors(mp):
    a, b = synth.small_pair(5000, seed=3)
    seed = mp.getSeed(11, 0)
    ok = np.array([[50, 1, 1], [50, 100, -200]], dtype=np.int64)
    lcb, ms = mp.libmems.anchor_scores(a, b, ok, [0, 1, 2], seed=seed)
    assert lcb.shape == (2,) and ms[0] == lcb[0]
    for bad in ([[50, 0, 1]], [[50, 1, len(b)]], [[-1, 1, 1]], [[len(a) + 1, 1, 1]]):
        with pytest.raises(mp.McuError):
            mp.libmems.anchor_scores(a, b, np.array(bad, dtype=np.int64), [0, 1], seed=seed)
    with pytest.raises(mp.McuError):
        mp.libmems.anchor_scores(a, b, ok, [0, 3], seed=seed)
    lcb, ms = mp.libmems.anchor_scores(a, b, np.zeros((0, 3), dtype=np.int64), [0, 0], seed=seed)   # an empty LCB scores 0
    assert lcb.tolist() == [0.0] and ms.size == 0


def test_sol_full_size_property(mp):
    """20 Mbp of the config-3 genome at its default weight: the frequencies equal an independent numpy evaluation (multiplicities
    from the device's own sorted list, exact integer window sums, one double division, one rounding to float)"""
    a, _b = synth.config3_pair(n=20_000_000)
    seq = a.tobytes()
    seed = mp.getSeed(mp.getDefaultSeedWeight(len(seq)), mp.CODING_SEED)
    L = mp.getSeedLength(seed)
    sml = mp.DNAMemorySML()
    sml.Create(seq, seed)
    sol = mp.SeedOccurrenceList()
    sol.construct(sml)
    f = sol.frequencies()
    want = P.sol_expected(sml.positions(), sml.mers(), sml.GetSeedMask(), len(seq), L)
    assert f.shape == want.shape and np.array_equal(f.view(np.uint32), want.view(np.uint32))
    assert f.min() >= 1.0 and (f != 1.0).sum() > 0


needs_next = pytest.mark.skipif(not os.path.exists(NEXT_BIN), reason="oracle/_ref/dropin_check_next not built (needs /root/reference at build time)")


@needs_next
def test_dropin_sol_and_scores_next_to_the_reference_classes(tmp_path):
    """the reference's SeedOccurrenceList / GetPairwiseAnchorScore and the C++ adapters in one process, on the LCBs the reference's
    own EliminateOverlaps_v2 / IdentifyBreakpoints / ComputeLCBs_v2 produce"""
    a, b = synth.small_pair(400000, seed=71, snp=0.02, n_inv=4)
    write_fasta(tmp_path / "a.fa", "a", a)
    write_fasta(tmp_path / "b.fa", "b", b)
    rc, kv, out = run_next(["sol", tmp_path / "a.fa", 15, 3])
    assert rc == 0 and kv["RESULT"] == "identical" and int(kv["not_one"]) > 0, out
    for w, r in ((15, 3), (11, 0), (0, 3)):
        rc, kv, out = run_next(["scores", tmp_path / "a.fa", tmp_path / "b.fa", w, r])
        assert rc == 0 and kv["RESULT"] == "identical" and int(kv["lcbs"]) >= 1, out


@needs_next
def test_dropin_scores_mds42(tmp_path):
    g0, g1 = _golden.mds42()
    write_fasta(tmp_path / "recoded.fa", "recoded", g0)
    write_fasta(tmp_path / "full.fa", "full", g1)
    rc, kv, out = run_next(["scores", tmp_path / "recoded.fa", tmp_path / "full.fa", 0, 3])
    assert rc == 0 and kv["RESULT"] == "identical" and kv["matches"] == "29403", out


# ---- gapped DP for regions with DNA wildcard columns (mcu_nw_batch_wild, csrc/dpwild.cu) -------------------------------------------
def test_nw_wild_golden_and_oracle(mp, orc):
    """paths of the reference's own NWSmall for inputs with N, X, R, Y, ... (golden), the float restatement on further inputs,
    and the integer kernel's paths on pure ACGT input"""
    z = _golden.npz("nw_wild.npz")
    n = int(z["n"])
    pairs = [(z["a%d" % i].tobytes(), z["b%d" % i].tobytes()) for i in range(n)]
    got = mp.GlobalAlignBatchWild(pairs)
    for i, p in enumerate(got):
        assert p.edges == z["p%d" % i].tobytes(), i
        assert p.score == _oracle.nw_align_f(*pairs[i])[1], i
    rng = np.random.default_rng(8)
    more = []
    for _ in range(300):
        wild = [b"N", b"NX", b"MRWSYKVHDBXNacgtn"][int(rng.integers(0, 3))]
        more.append((bytes(rng.choice(list(b"ACGT") * 5 + list(wild), int(rng.integers(1, 600))).astype(np.uint8)),
                     bytes(rng.choice(list(b"ACGT") * 5 + list(wild), int(rng.integers(1, 600))).astype(np.uint8))))
    more += [(b"N" * 3000, b"ACGT" * 700), (b"A", b"N"), (b"ACGTN" * 400, b"ACGTN" * 400)]
    for (a, b), p in zip(more, mp.GlobalAlignBatchWild(more)):
        want, score = _oracle.nw_align_f(a, b)
        assert p.edges == want and p.score == score, (len(a), len(b))
    acgt = synth.dp_pairs(60, 1, 800, seed=77)
    for p, q in zip(mp.GlobalAlignBatchWild(acgt), mp.GlobalAlignBatch(acgt)):
        assert p.edges == q.edges and p.score == float(q.score)


def test_nw_wild_errors(mp):
    from mauve_py_b200 import _capi
    with pytest.raises(mp.McuError) as e:
        mp.GlobalAlignBatchWild([(b"ACG1", b"ACGT")])
    assert e.value.code == _capi.MCU_EALPHA
    with pytest.raises(mp.McuError) as e:
        mp.GlobalAlignBatchWild([(b"ACG-", b"ACGT")])
    assert e.value.code == _capi.MCU_EALPHA
    big = (b"ACGTN" * 1000, b"A" * 2500 + b"ACNGT" * 500)   # 2.5e7 cells: beyond the old one-thread kernel's cap, 20 stripes of the wavefront
    (p,) = mp.GlobalAlignBatchWild([big])
    want, score = _oracle.nw_align_f(*big)
    assert p.edges == want and p.score == score
    with pytest.raises(mp.McuError):
        mp.GlobalAlignBatchWild([(b"", b"ACGT")])
    assert mp.GlobalAlignBatchWild([]) == []


def test_sml_accessors_on_a_device_built_list(mp, orc):
    """GetMer / GetSeedMer / FindMer of the mirror on a list built by mcu_sml_build: the mer at every sampled rank is
    GetSeedMer(position) (MemorySML::operator[], LM/MemorySML.cpp:88-94) and FindMer stops on a rank holding the query"""
    a, _ = synth.small_pair(40000, seed=8)
    for w, r in ((15, 3), (11, 0), (21, 0)):
        seed = mp.getSeed(w, r)
        sml = mp.DNAMemorySML()
        sml.Create(a, seed)
        rng = np.random.default_rng(w)
        for i in rng.integers(0, sml.SMLLength(), 200):
            b = sml[int(i)]
            assert sml.GetSeedMer(b.position) == b.mer
            found, at = sml.FindMer(b.mer)
            assert found and int(sml.mers()[at]) == b.mer
        assert sml.FindMer(int(sml.mers()[-1]) + 2)[0] is False or int(sml.mers()[-1]) + 2 in sml.mers()
        assert sml.Clone().SMLLength() == sml.SMLLength() and sml.GetHeader()["seed"] == seed


# ---- EliminateOverlaps_v2 + LengthFilter, IdentifyBreakpoints + ComputeLCBs_v2 (csrc/lcb.cu) --------------------------
def _lcb_cases(z):
    for nm in ["mds42", "syn"] + ["rand%d" % i for i in range(int(z["rand_count"]))]:
        rows = _golden.npz("mums_mds42.npz")["rows_w15_r3"] if nm == "mds42" else z[nm + "_rows"]
        yield nm, np.ascontiguousarray(rows, dtype=np.int64)


def test_eliminate_overlaps_and_lcbs_golden(mp):
    """the device pipeline against what the reference's own functions returned (tests/golden/lcb.npz): rows, their order, the LCB
    boundaries -- including the MDS42 list, whose second pass orders 25 tied rows the way libstdc++'s introsort leaves them"""
    z = _golden.npz("lcb.npz")
    for nm, rows in _lcb_cases(z):
        for key, both, ml in (("elim0", False, 0), ("elim0_min", False, 12), ("elim1", True, 0)):
            got, ties = mp.EliminateOverlaps_v2(rows, both, ml, return_ties=True)
            assert np.array_equal(got, z["%s_%s" % (nm, key)]), (nm, key, ties)
            if nm == "mds42" and key == "elim1":
                assert ties == 25
        e1 = z[nm + "_elim1"]
        if e1.shape[0]:
            so, bp = mp.IdentifyBreakpoints(e1)
            assert np.array_equal(so, z[nm + "_lcb_sorted"]) and np.array_equal(bp, z[nm + "_lcb_bp"]), nm
            lcbs = mp.ComputeLCBs_v2(so, bp)
            assert sum(len(l) for l in lcbs) == so.shape[0] and len(lcbs) == bp.size
    assert mp.EliminateOverlaps_v2(np.zeros((0, 3), dtype=np.int64)).shape == (0, 3)
    with pytest.raises(mp.McuError):
        mp.EliminateOverlaps_v2(np.array([[10, 0, 5]], dtype=np.int64))   # not a two-genome match


def test_eliminate_overlaps_and_lcbs_vs_oracle(mp, orc):
    """fresh random lists (dense overlaps on both strands, ties) and the match list of a real run: device == C restatement"""
    rng = np.random.default_rng(4242)
    lists = []
    for it in range(60):
        n = int(rng.integers(1, 3000))
        G = int(rng.integers(2000, 2000000))
        s0 = rng.integers(1, G, n)
        ln = rng.integers(5, 400, n)
        s1 = rng.integers(1, G, n) * rng.choice([1, -1], n)
        k = n // 2
        s0[:k] = np.sort(rng.integers(1, max(G // 10, 2), k))
        s1[:k] = s0[:k] + rng.integers(-3, 4, k)
        s1[s1 == 0] = 1
        lists.append(np.stack([ln, s0, s1], 1).astype(np.int64))
    a, b = synth.small_pair(2_000_000, seed=808, snp=0.02, n_inv=6)
    rows, _ = mp.libmems.find_mums(a, b, mp.getSeed(13, 0))
    lists.append(np.ascontiguousarray(rows, dtype=np.int64))
    for i, r in enumerate(lists):
        for both in (False, True):
            for ml in (0, 12):
                want, wt = _oracle.eliminate_overlaps(r, both, ml)
                got, gt = mp.EliminateOverlaps_v2(r, both, ml, return_ties=True)
                assert np.array_equal(got, want) and gt == wt, (i, both, ml)
        e1, _ = _oracle.eliminate_overlaps(r, True, 0)
        if e1.shape[0]:
            so, bp, t = _oracle.lcbs(e1)
            gso, gbp, gt = mp.IdentifyBreakpoints(e1, return_ties=True)
            assert np.array_equal(gso, so) and np.array_equal(gbp, bp) and gt == t, i


# ---- anchor columns of alignment windows (muscle::FindAnchorColsPP) ------------------------------------------------
def _bits(a):
    return np.asarray(a, dtype=np.float32).view(np.uint32)


def test_anchor_cols_golden(mp):
    """mcu_anchor_cols_batch against what the REFERENCE's FindAnchorColsPP / LetterObjScoreXP / WindowSmooth gave for the same windows
    (tests/golden/anchor_cols.npz): anchor columns, per-column scores and smoothed scores, float for float -- every window in one
    call (one CTA each), and each window on its own"""
    z = _golden.npz("anchor_cols.npz")
    n = int(z["n_windows"])
    wins = [(z["w%d_rows" % k], int(z["w%d_n1" % k]), z["w%d_weights" % k]) for k in range(n)]
    got = mp.libmems.FindAnchorColsPP_batch(wins, return_scores=True)
    total = 0
    for k in range(n):
        cols, score, smooth = got[k]
        assert np.array_equal(cols, z["w%d_cols" % k]), k
        assert np.array_equal(_bits(score), _bits(z["w%d_score" % k])) and np.array_equal(_bits(smooth), _bits(z["w%d_smooth" % k])), k
        total += cols.size
    assert total > 300
    for k in (0, 5, 13, 21, 29):
        rows, n1, w = wins[k]
        cols = mp.libmems.FindAnchorColsPP(rows[:n1], rows[n1:], weights=w)
        assert np.array_equal(cols, z["w%d_cols" % k]), k
    # two-genome windows need no weights; the default parameters are the settings the reference had in force
    assert np.array_equal(mp.libmems.FindAnchorColsPP(wins[13][0][:1], wins[13][0][1:]), z["w13_cols"])
    p = mp.libmems.AnchorParams.default()
    assert np.array_equal(_bits(np.array(p.subst[:])), _bits(z["settings"][:16])) and p.gap_open == float(z["settings"][16])
    assert mp.libmems.FindAnchorColsPP(b"ACGT" * 30, b"ACGT" * 31).size == 0      # different lengths: no anchor columns (MU/anchoredpp.cpp:358-362)


def test_anchor_cols_vs_oracle(mp):
    """fresh windows beside the oracle: 150 of the pipeline's two-row form and of alignments with more rows (random weights) in one
    batch, a window far longer than the kernel's tiles, other thresholds"""
    rng = np.random.default_rng(31)
    wins = []
    for it in range(150):
        ncol = int(rng.integers(1, 12000))
        n1, n2 = (1, 1) if it % 3 else (int(rng.integers(1, 4)), int(rng.integers(1, 4)))
        rows = synth.alignment_window(ncol, seed=5000 + it, n_rows=n1 + n2, snp=float(rng.choice([0.02, 0.1, 0.3])),
                                      gap_rate=float(rng.choice([0.002, 0.01, 0.06])), gap_mean=int(rng.choice([2, 12, 150])),
                                      both_gap=float(rng.choice([0.0, 0.002, 0.03])))
        wins.append((rows, n1, rng.random(n1 + n2).astype(np.float32) if it % 3 == 0 else None))
    wins.append((synth.alignment_window(1_000_000, seed=77, snp=0.05, gap_rate=0.005), 1, None))
    got = mp.libmems.FindAnchorColsPP_batch(wins, return_scores=True)
    n_cols = 0
    for k, (rows, n1, w) in enumerate(wins):
        c, s, m, _, _ = _oracle.anchor_cols(rows, n1, weights=w)
        assert np.array_equal(got[k][0], c), k
        assert np.array_equal(_bits(got[k][1]), _bits(s)) and np.array_equal(_bits(got[k][2]), _bits(m)), k
        n_cols += c.size
    assert n_cols > 5000
    # a batch beyond two windows per SM takes the other launch form (CTAs of 256 threads, per-column arrays in global memory)
    many = [(synth.alignment_window(int(rng.integers(30, 3000)), seed=9000 + it, gap_rate=float(rng.choice([0.002, 0.02])),
                                    gap_mean=int(rng.choice([4, 80]))), 1, None) for it in range(450)]
    got = mp.libmems.FindAnchorColsPP_batch(many, return_scores=True)
    for k, (rows, n1, w) in enumerate(many):
        c, s, m, _, _ = _oracle.anchor_cols(rows, n1)
        assert np.array_equal(got[k][0], c) and np.array_equal(_bits(got[k][1]), _bits(s)) and np.array_equal(_bits(got[k][2]), _bits(m)), k
    counters = np.zeros(8, dtype=np.uint64)
    mp.lib().mcu_test_anchor_counters(counters.ctypes.data)
    assert int(counters[0]) > 0 and int(counters[1]) > 0      # segments of the smoothing chain finished in exact arithmetic / as the float chain
    p = mp.libmems.AnchorParams.default()
    p.smooth_ceil, p.min_best_col, p.min_smooth, p.smooth_window, p.anchor_spacing, p.gap_extend = 120.0, 100.0, 60.0, 7, 32, -5.0
    po = _oracle.anchor_default_params()
    po.smooth_ceil, po.min_best_col, po.min_smooth, po.smooth_window, po.anchor_spacing, po.gap_extend = 120.0, 100.0, 60.0, 7, 32, -5.0
    for k in (1, 2, 3, 30, 31):
        rows, n1, w = wins[k]
        c, s, m, _, _ = _oracle.anchor_cols(rows, n1, weights=w, params=po)
        g = mp.libmems.FindAnchorColsPP_batch([(rows, n1, w)], params=p, return_scores=True)[0]
        assert np.array_equal(g[0], c) and np.array_equal(_bits(g[1]), _bits(s)) and np.array_equal(_bits(g[2]), _bits(m)), k


def test_anchor_cols_at_the_edges_of_the_exactness_tests(mp):
    """weights that push the scores to 1e8 and down to 1e-4 (sums that do / do not fit 24 bits, fractions of every size), rows of
    wildcards only, lengths around the smoothing window and around the tile size of the kernel, in one batch and one by one"""
    cases = []
    for scale in (1e6, 3.0e5, 1024.0, 0.5, 1.0 / 3.0, 1e-4):
        rows = synth.alignment_window(2600, seed=int(scale * 7) % 1000 + 1, snp=0.05, gap_rate=0.003)
        cases.append((rows, 1, np.array([scale, 1.0], dtype=np.float32)))
        cases.append((rows, 1, np.array([scale, scale], dtype=np.float32)))
    for ncol in (21, 22, 23, 42, 43, 511 + 21, 512 + 21, 513 + 21, 1023 + 21, 1024 + 21, 1025 + 21, 2048 + 21, 2049 + 21):
        cases.append((synth.alignment_window(ncol, seed=ncol, gap_rate=0.004), 1, np.ones(2, dtype=np.float32)))
    wild = np.full((2, 900), ord("N"), dtype=np.uint8)
    cases.append((wild, 1, np.ones(2, dtype=np.float32)))
    half = synth.alignment_window(900, seed=3)
    half[1, :450] = ord("N")
    cases.append((half, 1, np.ones(2, dtype=np.float32)))
    batch = mp.libmems.FindAnchorColsPP_batch(cases, return_scores=True)
    for k, (rows, n1, w) in enumerate(cases):
        c, s, m, _, _ = _oracle.anchor_cols(rows, n1, weights=w)
        one = mp.libmems.FindAnchorColsPP_batch([(rows, n1, w)], return_scores=True)[0]
        for got in (batch[k], one):
            assert np.array_equal(got[0], c) and np.array_equal(_bits(got[1]), _bits(s)) and np.array_equal(_bits(got[2]), _bits(m)), (k, rows.shape, w)


def test_anchor_cols_argument_errors(mp):
    rows = synth.alignment_window(100, seed=1)
    p = mp.libmems.AnchorParams.default()
    p.smooth_window = 20                      # WindowSmooth quits on an even window (MU/anchors.cpp:14-15)
    with pytest.raises(mp.McuError):
        mp.libmems.FindAnchorColsPP_batch([(rows, 1)], params=p)
    with pytest.raises(ValueError):
        mp.libmems.FindAnchorColsPP_batch([(rows, 2)])      # an alignment without rows
    assert mp.libmems.FindAnchorColsPP_batch([]) == []
