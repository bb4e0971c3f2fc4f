"""GPU (-m gpu): the CUDA path through the C ABI against the oracle on the same inputs and against the golden vectors."""
import hashlib

import numpy as np
import pytest

import _golden
from mauve_py_b200 import synth

pytestmark = pytest.mark.gpu


# ---- radix sort (test hook) ---------------------------------------------------------------------
@pytest.mark.parametrize("dtype,bits", [(np.uint32, 32), (np.uint32, 13), (np.uint64, 64), (np.uint64, 40), (np.uint64, 8)])
@pytest.mark.parametrize("n", [0, 1, 31, 8191, 8192, 8193, 100003, 1 << 20])
def test_radix_sort_pairs(mp, dtype, bits, n):
    rng = np.random.default_rng(n + bits)
    hi = (1 << bits) - 1
    keys = rng.integers(0, hi, n, dtype=np.uint64, endpoint=True).astype(dtype)
    if n > 100:
        keys[: n // 3] = keys[0]  # heavy duplicates: stability matters
    vals = np.arange(n, dtype=np.uint32)
    k2, v2 = mp.sort_pairs(keys, vals, bits)
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(k2, keys[order])
    assert np.array_equal(v2, vals[order])


# ---- sorted mer list ----------------------------------------------------------------------------
def test_sml_golden(mp):
    z = _golden.npz("sml_small.npz")
    for name, w, r, seed in _golden.cases(z):
        sml = mp.DNAMemorySML()
        sml.Create(z["seq_" + name].tobytes(), seed)
        key = "%s_w%d_r%d" % (name, w, r)
        assert np.array_equal(sml.mers(), z["mer_" + key]), key
        assert np.array_equal(sml.positions(), z["pos_" + key]), key
        assert sml.SMLLength() == z["mer_" + key].size and sml.Seed() == seed


@pytest.mark.parametrize("n,w,r", [(20, 5, 0), (21, 15, 3), (1000, 7, 1), (65537, 11, 0), (300000, 15, 3), (300000, 16, 2), (200001, 19, 0),
                                   (150000, 21, 0), (100000, 24, 0), (100000, 31, 0), (50000, 3, 0)])
def test_sml_vs_oracle(mp, orc, n, w, r):
    seq = synth.random_genome(n, 0.45, synth.rng_for(n + w)).tobytes()
    seed = mp.getSeed(w, r)
    sml = mp.DNAMemorySML()
    sml.Create(seq, seed)
    opos, omer = orc.sml_build(seq, seed)
    assert np.array_equal(sml.mers(), omer)
    assert np.array_equal(sml.positions(), opos)
    # packed sequence = SortedMerList::sequence
    import _oracle
    words = int(_oracle.oracle().orc_packed_words(n))
    ref_packed = np.zeros(words, dtype=np.uint32)
    _oracle.oracle().orc_pack(seq, n, ref_packed.ctypes.data)
    assert np.array_equal(sml.packed_sequence(), ref_packed)


def test_sml_edge_cases(mp, orc):
    seed = mp.getSeed(5, 0)
    L = mp.getSeedLength(seed)
    for seq in [b"", b"A", b"ACGT"[: L - 1], b"ACGTACGTAC"[:L], b"acgtnnRYKMSWBDHVxyz!" * 3]:
        sml = mp.DNAMemorySML()
        sml.Create(seq, seed)
        opos, omer = orc.sml_build(seq, seed)
        assert sml.SMLLength() == opos.size
        assert np.array_equal(sml.mers(), omer) and np.array_equal(sml.positions(), opos)
    with pytest.raises(mp.McuError) as e:  # '-' throws in the reference (LM/SortedMerList.cpp:433-437)
        mp.DNAMemorySML().Create(b"ACGTACGT-ACGTACGTACGTACGT", seed)
    from mauve_py_b200 import _capi
    assert e.value.code == _capi.MCU_EGAP
    with pytest.raises(mp.McuError):  # pattern with a trailing zero is not a DNA seed
        mp.DNAMemorySML().Create(b"ACGTACGTACGTACGTACGT", 0b10110)


def test_sml_mds42_digest(mp):
    m = _golden.meta(_golden.npz("mums_mds42.npz"))["sml_full_w15_r3"]
    _, g1 = _golden.mds42()
    sml = mp.DNAMemorySML()
    sml.Create(g1, 0x16df6d)
    assert sml.SMLLength() == m["n"]
    assert hashlib.sha1(sml.mers().tobytes()).hexdigest() == m["sha1_mer"]
    assert hashlib.sha1(sml.positions().tobytes()).hexdigest() == m["sha1_pos_canon"]
    found, idx = sml.FindMer(int(sml.mers()[12345]))
    assert found and (sml.mers()[idx] >> np.uint64(34)) == (sml.mers()[12345] >> np.uint64(34))


# ---- seed + match + extend ---------------------------------------------------------------------
def test_mums_golden_small(mp):
    z = _golden.npz("mums_small.npz")
    for i, w, r, rule, seed, coll, cnt in _golden.cases(z):
        a, b = z["a%d" % i].tobytes(), z["b%d" % i].tobytes()
        ml = mp.MatchList(seq_table=[a, b])
        ml.CreateMemorySMLs(w, r)
        mh = mp.PairwiseMatchFinder() if rule == 0 else mp.MemHash()
        mh.FindMatches(ml)
        assert np.array_equal(ml.as_array(), z["rows%d" % i]), i
        assert mh.MemCount() == cnt and mh.MemCollisionCount() == coll, i


def test_mums_mds42(mp):
    z = _golden.npz("mums_mds42.npz")
    m = _golden.meta(z)
    g0, g1 = _golden.mds42()
    rows, stats = mp.libmems.find_mums(g0, g1, m["w15_r3"]["seed"])
    assert rows.shape[0] == 29403 and int(rows[:, 0].sum()) == 3792460 and int((rows[:, 2] < 0).sum()) == 1515
    assert np.array_equal(rows, z["rows_w15_r3"])
    assert int(stats[0]) == 2744091 and int(stats[2]) == m["w15_r3"]["collisions"]
    for key in ("w15_r0", "w11_r0", "w21_r0"):
        rows, _ = mp.libmems.find_mums(g0, g1, m[key]["seed"])
        assert rows.shape[0] == m[key]["n"] and int(rows[:, 0].sum()) == m[key]["sum_len"]
        assert hashlib.sha1(rows.tobytes()).hexdigest() == m[key]["sha1_rows"], key


@pytest.mark.parametrize("n,w,r,kw", [(200000, 11, 0, {}), (400000, 15, 3, dict(snp=0.01, n_inv=4)), (300000, 13, 1, dict(snp=0.08)),
                                      (250000, 19, 0, dict(snp=0.003, n_inv=2)), (100000, 9, 0, dict(snp=0.02)),
                                      (120000, 24, 0, dict(snp=0.001))])
def test_mums_vs_oracle(mp, orc, n, w, r, kw):
    a, b = synth.small_pair(n, seed=n + w, **kw)
    seed = mp.getSeed(w, r)
    rows, stats = mp.libmems.find_mums(a, b, seed)
    orows, ostats = orc.find_mums(a, b, seed, 0)
    assert np.array_equal(rows, orows)
    assert int(stats[0]) == int(ostats[3]) and int(stats[1]) == int(ostats[1])


def test_mums_config2_slice_vs_oracle(mp, orc):
    a, b = synth.config2_pair(n=1_000_000)
    seed = mp.getSeed(15, mp.CODING_SEED)
    rows, stats = mp.libmems.find_mums(a, b, seed)
    orows, _ = orc.find_mums(a.tobytes(), b.tobytes(), seed, 0)
    assert rows.shape[0] > 100 and np.array_equal(rows, orows)


def test_mums_self_and_revcomp(mp, orc):
    g = synth.random_genome(150000, 0.5, synth.rng_for(9))
    seed = mp.getSeed(13, 0)
    for other in (g, synth.revcomp(g)):
        rows, _ = mp.libmems.find_mums(g, other, seed)
        orows, _ = orc.find_mums(g.tobytes(), other.tobytes(), seed, 0)
        assert np.array_equal(rows, orows) and rows.shape[0] >= 1


def test_mums_sharded_equals_unsharded(mp):
    a, b = synth.small_pair(500000, seed=77, snp=0.01, n_inv=3)
    seed = mp.getSeed(15, 3)
    full, _ = mp.libmems.find_mums(a, b, seed)
    for world in (2, 3, 8):
        parts = []
        s = mp.AnchorSession()
        s.upload(a, b)
        for rank in range(world):
            s.run(seed, rank, world)
            parts.append(s.download().copy())
        s.close()
        merged = mp.merge_matches(np.concatenate(parts, axis=0))
        assert np.array_equal(merged, full), world


def test_session_matches_one_shot(mp):
    a, b = synth.small_pair(300000, seed=3)
    seed = mp.getSeed(15, 3)
    rows, stats = mp.libmems.find_mums(a, b, seed)
    s = mp.AnchorSession()
    s.upload(a, b)
    for _ in range(3):  # re-running a session is idempotent
        n = s.run(seed)
        assert n == rows.shape[0] and np.array_equal(s.download(), rows)
    assert s.launch_count() > 0 and s.stage_ms[6] > 0
    s.close()


# ---- gapped DP ------------------------------------------------------------------------------------
def test_nw_golden(mp):
    z = _golden.npz("nw_small.npz")
    n = int(z["n"])
    pairs = [(z["a%d" % i].tobytes(), z["b%d" % i].tobytes()) for i in range(n)]
    paths = mp.GlobalAlignBatch(pairs)
    for i, p in enumerate(paths):
        assert p.edges == z["p%d" % i].tobytes(), i


def test_nw_vs_oracle(mp, orc):
    pairs = synth.dp_pairs(120, 1, 3000, seed=31)
    rng = synth.rng_for(5)
    for la, lb in [(255, 257), (256, 256), (257, 255), (512, 1), (1, 512), (513, 700), (2049, 2047), (5000, 4800), (300, 6000)]:
        pairs.append((synth.random_genome(la, 0.5, rng).tobytes(), synth.random_genome(lb, 0.5, rng).tobytes()))
    paths = mp.GlobalAlignBatch(pairs)
    for (a, b), p in zip(pairs, paths):
        op, osc = orc.nw_align(a, b)
        assert p.score == osc, (len(a), len(b))
        assert p.edges == op, (len(a), len(b))


def test_nw_path_invariants_large(mp):
    """size-independent properties at the window cap (20 kbp): path consumes both sequences, score recomputes from the path"""
    rng = synth.rng_for(41)
    a = synth.random_genome(20000, 0.5, rng)
    b = synth.indels(synth.snps(a, 0.05, rng), 200, 0.3, 30, rng)
    p = mp.GlobalAlign(a.tobytes(), b.tobytes())
    e = np.frombuffer(p.edges, dtype=np.uint8)
    assert int(((e == ord("M")) | (e == ord("D"))).sum()) == a.size
    assert int(((e == ord("M")) | (e == ord("I"))).sum()) == b.size
    # recompute the score of the returned path with the restated scoring (S + gaps: -400 interior, -200 terminal)
    NUC = np.array([[151, -54, 29, -63], [-54, 160, -65, 29], [29, -65, 160, -54], [-63, 29, -54, 151]])
    code = synth._CODE
    ia = np.cumsum((e == ord("M")) | (e == ord("D"))) - 1
    ib = np.cumsum((e == ord("M")) | (e == ord("I"))) - 1
    m = e == ord("M")
    score = int(NUC[code[a[ia[m]]], code[b[ib[m]]]].sum())
    starts = np.flatnonzero((e != ord("M")) & (np.concatenate(([ord("M")], e[:-1])) != e))
    for s in starts:
        t = s
        while t < e.size and e[t] == e[s]:
            t += 1
        score -= 200 if (s == 0 or t == e.size) else 400
    assert score == p.score


def test_nw_errors(mp):
    from mauve_py_b200 import _capi
    with pytest.raises(mp.McuError) as e:
        mp.GlobalAlign(b"ACGTN", b"ACGT")
    assert e.value.code == _capi.MCU_EALPHA
    with pytest.raises(mp.McuError):
        mp.GlobalAlign(b"", b"ACGT")
    assert mp.GlobalAlignBatch([]) == []


def test_sml_shards_concatenate_to_the_sorted_list(mp):
    """SURVEY 8e, sorted mer list sharded by mer range: the shards' lists, one after the other, carry the mer sequence of the
    unsharded list and, run by run, the same positions; the ranges are balanced"""
    g = synth.random_genome(3_000_000, 0.47, synth.rng_for(17)).tobytes()
    for w, r in ((15, 3), (19, 3), (11, 0), (21, 0)):
        seed = mp.getSeed(w, r)
        sml = mp.DNAMemorySML()
        sml.Create(g, seed)
        pos_u, mer_u = sml.positions(), sml.mers()
        for world in (2, 5, 8):
            parts = [mp.libmems.sml_build_shard(g, seed, k, world) for k in range(world)]
            pos_s = np.concatenate([p for p, _ in parts])
            mer_s = np.concatenate([m for _, m in parts])
            assert np.array_equal(mer_s, mer_u), (w, r, world)
            assert np.array_equal(np.sort(pos_s), np.arange(pos_u.size, dtype=np.uint32))
            # run by run the same positions: sorting (mer, position) pairs makes the order inside runs comparable
            a = np.lexsort((pos_s, mer_s))
            b = np.lexsort((pos_u, mer_u))
            assert np.array_equal(pos_s[a], pos_u[b]), (w, r, world)
            sizes = np.array([p.size for p, _ in parts], dtype=np.float64)
            assert sizes.max() / sizes.mean() < 1.25, (w, r, world, sizes.tolist())
    with pytest.raises(mp.McuError):
        mp.libmems.sml_build_shard(g, mp.getSeed(15, 3), 3, 3)


# ---- HMM ------------------------------------------------------------------------------------------
def _check_hmm(pred, post, ref_pred, ref_post, exact=True):
    """north_star bar: 1e-5 relative on the posterior.  The default (bfloat-faithful) path is held to more: the reference's
    posteriors bit for bit (a last-ulp difference of the device's exp/division would show up as < 1e-12) and identical calls."""
    rel = float(np.max(np.abs(post - ref_post) / np.maximum(ref_post, 1e-300))) if len(post) else 0.0
    if exact:
        assert rel < 1e-12, rel
        assert pred == ref_pred
        return
    assert np.allclose(post, ref_post, rtol=1e-5, atol=1e-30), rel
    mism = np.flatnonzero(np.frombuffer(pred, dtype=np.uint8) != np.frombuffer(ref_pred, dtype=np.uint8))
    assert all(abs(ref_post[j] - 0.9) <= 1e-5 for j in mism)


def test_hmm_golden(mp):
    z = _golden.npz("hmm_small.npz")
    for c in _golden.cases(z):
        i = c[0]
        pred, post = mp.run(z["sym%d" % i].tobytes(), z["params%d" % i], want_posterior=True)
        _check_hmm(pred, post, z["pred%d" % i].tobytes(), z["post%d" % i])


def test_hmm_batch_vs_oracle(mp, orc):
    params = mp.libmems.hmm_params(0.5, 1e-5, 1e-9, 0.7)
    seqs = [synth.hmm_string(n, seed=n, block=b) for n, b in [(1, 10), (2, 10), (63, 40), (64, 40), (65, 40), (128, 50), (1000, 100), (100000, 500),
                                                              (1_000_000, 3000)]]
    seqs.insert(3, b"")
    preds, posts, ms = mp.run_batch(seqs, params, want_posterior=True)
    for s, pred, post in zip(seqs, preds, posts):
        if len(s) == 0:
            assert pred == b""
            continue
        opred, opost = orc.hmm_run(s, params)
        _check_hmm(pred, post, opred, opost)
    from mauve_py_b200 import _capi
    with pytest.raises(mp.McuError) as e:
        mp.run(b"12349", params)
    assert e.value.code == _capi.MCU_EINVAL


def test_hmm_long_string_fp32_chain(mp, orc, monkeypatch):
    """ONE long string (what progressiveMauve really asks for): the warp chain's FP32 form of the recurrence (csrc/hmm.cu) gives the
    posteriors of the operation-by-operation FP64 form bit for bit, takes the FP64 form at a handful of hazardous columns only, and
    both equal the reference's run() on a prefix the oracle finishes quickly"""
    import ctypes as C
    params = mp.libmems.hmm_params(0.5, 0.0, 0.0, 0.0)
    n = 2_000_000
    s = synth.hmm_string(n, seed=77, block=2500)
    c3 = np.zeros(3, dtype=np.uint64)
    preds, posts, ms_fast = mp.run_batch([s], params, want_posterior=True)
    assert mp.lib().mcu_test_hmm_counters(c3.ctypes.data) == 0
    assert int(c3[0]) == 2 * (n - 1)
    assert 0 < int(c3[2]) < n // 200           # hazardous columns exist and are rare
    assert int(c3[1]) < n // 4                 # chain rounds: far fewer than columns
    monkeypatch.setenv("MAUVE_CUDA_HMM_FP64", "1")
    preds64, posts64, ms_64 = mp.run_batch([s], params, want_posterior=True)
    assert mp.lib().mcu_test_hmm_counters(c3.ctypes.data) == 0
    assert int(c3[2]) == 2 * (n - 1)
    monkeypatch.delenv("MAUVE_CUDA_HMM_FP64")
    assert preds[0] == preds64[0]
    assert np.array_equal(posts[0].view(np.uint64), posts64[0].view(np.uint64))
    print("hmm 2M columns: FP32 chain %.1f ms, FP64 chain %.1f ms" % (ms_fast, ms_64))
    m = 300_000
    preds, posts, _ = mp.run_batch([s[:m]], params, want_posterior=True)
    opred, opost = orc.hmm_run(s[:m], params)
    _check_hmm(preds[0], posts[0], opred, opost)
    assert np.array_equal(posts[0].view(np.uint64), np.asarray(opost).view(np.uint64))


@pytest.mark.parametrize("every", [1, 3, 37, 1000])
def test_hmm_chain_repair_protocol(mp, orc, monkeypatch, every):
    """the long-string kernel repairs its chain when a parked value is not the reference's (a rounding hazard that mattered: about
    one column in 20 million, so real inputs hardly ever get there).  MAUVE_CUDA_HMM_TEST_FAULT=N flips the last mantissa bit of the
    chain's state at the end of every N-th group of eight columns -- in the parked value and in the chain itself; the re-examination
    has to notice, correct the column, have the block redone behind it and the chain restarted: posteriors still bit-identical"""
    params = mp.libmems.hmm_params(0.5, 0.0, 0.0, 0.0)
    n = 200_003
    s = synth.hmm_string(n, seed=5150 + every, block=700)
    opred, opost = orc.hmm_run(s, params)
    monkeypatch.setenv("MAUVE_CUDA_HMM_TEST_FAULT", str(every))
    preds, posts, _ = mp.run_batch([s], params, want_posterior=True)
    c3 = np.zeros(3, dtype=np.uint64)
    assert mp.lib().mcu_test_hmm_counters(c3.ctypes.data) == 0
    monkeypatch.delenv("MAUVE_CUDA_HMM_TEST_FAULT")
    assert preds[0] == opred
    assert np.array_equal(posts[0].view(np.uint64), np.asarray(opost).view(np.uint64))
    blocks = 2 * ((n - 1 + 31) // 32)
    assert int(c3[1]) > blocks            # re-examination rounds: one per block plus the repairs


def test_hmm_scan_mode_within_tolerance(mp, orc, monkeypatch):
    """MAUVE_CUDA_HMM_SCAN=1: the column-parallel evaluation in double stays within the 1e-5 bar for strings up to ~10 k columns"""
    monkeypatch.setenv("MAUVE_CUDA_HMM_SCAN", "1")
    params = mp.libmems.hmm_params(0.45, 1e-5, 1e-9, 0.7)
    seqs = [synth.hmm_string(n, seed=n + 1, block=b) for n, b in [(1, 10), (64, 40), (65, 40), (1000, 100), (10000, 300)]]
    preds, posts, ms = mp.run_batch(seqs, params, want_posterior=True)
    for s, pred, post in zip(seqs, preds, posts):
        opred, opost = orc.hmm_run(s, params)
        _check_hmm(pred, post, opred, opost, exact=False)


@pytest.mark.parametrize("w,sd", [(11, 5), (15, 6), (9, 7), (13, 8)])
def test_mums_order_dependent_buckets(mp, orc, w, sd):
    """diagonals that collide mod 40000 with interleaved spans: the reference stores some matches twice (csrc/replay.cu)"""
    a, b = synth.colliding_diagonals_pair(seed=sd)
    seed = mp.getSeed(w, 0)
    rows, stats = mp.libmems.find_mums(a, b, seed)
    orows, ostats = orc.find_mums(a, b, seed, 0)
    assert orows.shape[0] > np.unique(orows, axis=0).shape[0]  # the reference really duplicates rows here
    assert int(stats[6]) > 0 and int(stats[7]) == orows.shape[0] - np.unique(orows, axis=0).shape[0]
    assert np.array_equal(rows, orows)
    assert int(stats[1]) == int(ostats[1]) and int(stats[2]) == int(ostats[0])


def test_merge_reports_order_dependent_buckets(mp):
    a, b = synth.colliding_diagonals_pair(seed=5)
    seed = mp.getSeed(11, 0)
    s = mp.AnchorSession()
    s.upload(a, b)
    parts = []
    for rank in range(2):
        s.run(seed, rank, 2)
        parts.append(s.download().copy())
    s.close()
    merged, unclean = mp.merge_matches(np.concatenate(parts, axis=0), return_unclean=True)
    assert unclean > 0
    full, _ = mp.libmems.find_mums(a, b, seed)
    assert np.array_equal(np.unique(merged, axis=0), np.unique(full, axis=0))
    a2, b2 = synth.small_pair(200000, seed=3)
    s = mp.AnchorSession()
    s.upload(a2, b2)
    s.run(seed)
    merged, unclean = mp.merge_matches(s.download(), return_unclean=True)
    assert unclean == 0
    s.close()


@pytest.mark.parametrize("w", [19, 17, 21, 23, 31])
def test_mums_solid_seed_paths(mp, orc, w):
    """solid odd-weight seeds take the run-length extension (extend_solid_kernel): same rows as the oracle, incl. sequence ends"""
    a, b = synth.small_pair(300000, seed=w, snp=0.01, n_inv=3)
    # matches touching both ends of both genomes, forward and reverse
    tail = synth.random_genome(400, 0.5, synth.rng_for(w + 1)).tobytes()
    a2 = tail + a + synth.revcomp(np.frombuffer(tail, dtype=np.uint8)).tobytes()
    b2 = tail + b + synth.revcomp(np.frombuffer(tail, dtype=np.uint8)).tobytes()
    seed = mp.getSeed(w, mp.SOLID_SEED)
    assert mp.getSeedLength(seed) == w
    for x, y in ((a, b), (a2, b2), (a2, synth.revcomp(np.frombuffer(b2, dtype=np.uint8)).tobytes())):
        rows, stats = mp.libmems.find_mums(x, y, seed)
        orows, _ = orc.find_mums(x, y, seed, 0)
        assert rows.shape[0] > 10 and np.array_equal(rows, orows)


# ---- two-phase sharded run (enumerate | all-reduce of the unique-seed bitmap | finish | merge) ---------------
def _two_phase(mp, a, b, seed, world):
    """world sessions on one GPU stand in for world ranks; the bitmap SUM is what comm.cu's ncclAllReduce does between real ranks"""
    import torch
    from _devwords import DeviceWords as _DeviceWords
    sess = []
    for rank in range(world):
        s = mp.AnchorSession()
        s.upload(a, b)
        s.enumerate(seed, rank, world)
        sess.append(s)
    views = [torch.as_tensor(_DeviceWords(*s.uniq_bitmap()), device="cuda") for s in sess]
    total = torch.stack(views).sum(dim=0, dtype=torch.int32)
    ored = views[0].clone()
    for v in views[1:]:
        ored |= v
    assert torch.equal(total, ored)  # the slices' bits are disjoint: SUM == OR
    for v in views:
        v.copy_(total)
    torch.cuda.synchronize()
    parts, emitted = [], 0
    for s in sess:
        n = s.finish(uniq_is_global=True)
        emitted += n
        parts.append(s.download().copy())
    rows = np.concatenate(parts, axis=0)
    n, st = sess[0].merge(rows)
    out = sess[0].download().copy()
    for s in sess:
        s.close()
    return out, emitted, st


@pytest.mark.parametrize("w,r,n", [(15, 3, 500000), (19, 3, 400000), (11, 0, 200000), (21, 0, 300000)])
def test_two_phase_sharded_exact(mp, w, r, n):
    a, b = synth.small_pair(n, seed=100 + w, snp=0.01, n_inv=3)
    seed = mp.getSeed(w, r)
    full, _ = mp.libmems.find_mums(a, b, seed)
    for world in (2, 3, 8):
        out, emitted, st = _two_phase(mp, a, b, seed, world)
        assert emitted == full.shape[0], (world, emitted, full.shape[0])  # exactly one emitter per match across ranks
        assert np.array_equal(out, full), world


def test_two_phase_sharded_order_dependent_buckets(mp, orc):
    """the rank-0 merge replays the reference's order-dependent hash buckets with the global bitmap: duplicate rows included"""
    a, b = synth.colliding_diagonals_pair(seed=5)
    seed = mp.getSeed(11, 0)
    orows, _ = orc.find_mums(a, b, seed, 0)
    out, emitted, st = _two_phase(mp, a, b, seed, 2)
    assert int(st[0]) > 0 and np.array_equal(out, orows)


# ---- bucketed enumeration: repeats overflow the fixed-capacity buckets ------------------------------------------
@pytest.mark.parametrize("w,r", [(15, 3), (19, 3), (11, 0)])
def test_mums_repeat_rich_overflow_paths(mp, orc, monkeypatch, w, r):
    """many copies of one mer: overflow records + dirty buckets take the radix-sort path; a full overflow array makes the
    run start over with exact bucket sizes; all three give the oracle's rows"""
    a, b = synth.repeat_rich_pair(seed=w)
    seed = mp.getSeed(w, r)
    orows, ostats = orc.find_mums(a, b, seed, 0)
    s = mp.AnchorSession()
    s.upload(a, b)
    s.run(seed)
    assert s.stage_ms[7] == -1 and s.stage_ms[15] > 0       # fixed-capacity layout, some buckets spilled to the sort path
    assert int(s.stats[3]) == 1                              # a run longer than MER_REPEAT_LIMIT was seen
    assert np.array_equal(s.download(), orows)
    monkeypatch.setenv("MAUVE_CUDA_OVF_CAP", "100")
    s.run(seed)
    assert s.stage_ms[7] == -2                               # fell back to exact bucket sizes
    assert np.array_equal(s.download(), orows)
    monkeypatch.delenv("MAUVE_CUDA_OVF_CAP")
    monkeypatch.setenv("MAUVE_CUDA_EXACT_BUCKETS", "1")
    s.run(seed)
    assert s.stage_ms[7] == -2 and np.array_equal(s.download(), orows)
    monkeypatch.delenv("MAUVE_CUDA_EXACT_BUCKETS")
    for world in (2, 5):                                     # sharded: fixed level-1 bin ranges per rank
        parts = []
        for rank in range(world):
            s.run(seed, rank, world)
            parts.append(s.download().copy())
        assert np.array_equal(np.unique(mp.merge_matches(np.concatenate(parts, axis=0)), axis=0), np.unique(orows, axis=0))
    s.close()


# ---- batched gap search (recursive anchoring: many small pairs per call) ------------------------------------------
def _gap_pairs(n, lo, hi, seed):
    rng = np.random.default_rng(seed)
    pairs = []
    for i in range(n):
        la = int(np.exp(rng.uniform(np.log(lo), np.log(hi))))
        a, b = synth.small_pair(la, seed=seed * 1000 + i, snp=0.03, n_inv=1 if i % 3 == 0 else 0)
        if i % 5 == 0:
            b = synth.revcomp(np.frombuffer(b, dtype=np.uint8)).tobytes()
        if i % 7 == 0:
            b = b[: len(b) // 2]
        pairs.append((a, b))
    return pairs


def test_gap_batch_vs_oracle(mp, orc):
    """every pair of the batch gets the rows the oracle (= MemHash on that pair alone) gives, in the same order"""
    pairs = _gap_pairs(300, 30, 6000, 5) + [(b"", b"ACGT"), (b"ACGTACGTAC", b"ACGTACGTAC"), (b"A" * 500, b"A" * 400)]
    res, stats = mp.libmems.find_mums_batch(pairs)
    assert len(res) == len(pairs) and int(stats[3]) >= 3          # several seed weights in one call
    nonempty = 0
    for (a, b), rows in zip(pairs, res):
        w = mp.getDefaultSeedWeight((len(a) + len(b)) // 2)
        if w < 5:
            assert rows.shape[0] == 0
            continue
        orows, _ = orc.find_mums(a, b, mp.getSeed(w, 0), 1)
        assert np.array_equal(rows, orows), (len(a), len(b), w)
        nonempty += rows.shape[0] > 0
    assert nonempty > 200


def test_gap_batch_equals_single_pair_calls(mp):
    """same rows as mcu_find_mums pair by pair, including a pair with order-dependent hash buckets (redone one by one)"""
    pairs = _gap_pairs(40, 200, 20000, 9)
    pairs.insert(7, synth.colliding_diagonals_pair(seed=5))
    seeds = [mp.getSeed(11, 0)] * len(pairs)
    res, stats = mp.libmems.find_mums_batch(pairs, seeds=seeds)
    assert int(stats[2]) >= 1 and int(stats[3]) == 1
    for (a, b), rows in zip(pairs, res):
        single, _ = mp.libmems.find_mums(a, b, seeds[0], 1)
        assert np.array_equal(rows, single)


# ---- chunked upload: pack + level-1 partition of a piece run under the copies of the next pieces (mcu_session_upload_begin) ----
@pytest.mark.parametrize("w,r,n,chunks", [(15, 3, 400000, 1), (15, 3, 400000, 3), (19, 3, 350000, 8), (13, 0, 277777, 5), (19, 3, 70000, 64),
                                          (9, 0, 20000, 4)])
def test_chunked_upload_overlap_exact(mp, orc, w, r, n, chunks):
    """whatever the piece boundaries (a piece is a whole number of 4096-position scatter tiles; seeds reach into the next piece;
    genome 0's last tile straddles into genome 1), the rows equal the oracle's; sizes below the bucketed plan take the unoverlapped path"""
    a, b = synth.small_pair(n, seed=7 * n + w, snp=0.01, n_inv=2)
    seed = mp.getSeed(w, r)
    ha, hb = np.frombuffer(a, dtype=np.uint8).copy(), np.frombuffer(b, dtype=np.uint8).copy()
    s = mp.AnchorSession()
    for _ in range(2):   # twice: the second upload starts while the first run's buffers are the live ones
        s.upload_begin_ptr(ha.ctypes.data, ha.size, hb.ctypes.data, hb.size, chunks)
        cnt = s.run(seed)
        rows = s.download().copy()
        orows, _ = orc.find_mums(a, b, seed, 0)
        assert cnt == orows.shape[0] and np.array_equal(rows, orows)
    # a plain upload after a chunked one leaves no stale piece state behind
    s.upload(a, b)
    assert s.run(seed) == orows.shape[0] and np.array_equal(s.download(), orows)
    s.close()


def test_find_mums_into_caller_buffer(mp, orc):
    """mcu_find_mums_into: rows land in caller memory; a buffer that is too small is refused with the required size"""
    import ctypes as C
    a, b = synth.small_pair(300000, seed=5, snp=0.01)
    seed = mp.getSeed(15, 3)
    orows, _ = orc.find_mums(a, b, seed, 0)
    buf = np.zeros((orows.shape[0] + 5, 3), dtype=np.int64)
    n_out = C.c_uint64(0)
    stats = np.zeros(8, dtype=np.uint64)
    rc = mp.lib().mcu_find_mums_into(a, len(a), b, len(b), seed, 0, buf.ctypes.data, buf.shape[0], C.byref(n_out), stats.ctypes.data)
    assert rc == 0 and n_out.value == orows.shape[0] and np.array_equal(buf[:n_out.value], orows)
    rc = mp.lib().mcu_find_mums_into(a, len(a), b, len(b), seed, 0, buf.ctypes.data, 3, C.byref(n_out), None)
    assert rc == -7 and n_out.value == orows.shape[0]


def test_sharded_entry_points_with_one_rank(mp, orc):
    """a one-rank communicator makes the collective entry points plain calls (no NCCL): same rows"""
    import ctypes as C
    lib = mp.lib()
    if lib.mcu_comm_world() == 1 and lib.mcu_comm_rank() == 0:
        lib.mcu_comm_init(0, 1, None)    # idempotence is not promised: a second call answers MCU_EINVAL and changes nothing
    a, b = synth.small_pair(200000, seed=9, snp=0.01)
    seed = mp.getSeed(13, 0)
    orows, _ = orc.find_mums(a, b, seed, 0)
    s = mp.AnchorSession()
    s.upload(a, b)
    assert s.run_sharded(seed) == orows.shape[0] and np.array_equal(s.download(), orows)
    buf = np.zeros((orows.shape[0] + 1, 3), dtype=np.int64)
    n_out = C.c_uint64(0)
    assert lib.mcu_find_mums_sharded(a, len(a), b, len(b), seed, 0, buf.ctypes.data, buf.shape[0], C.byref(n_out), None) == 0
    assert np.array_equal(buf[:n_out.value], orows)
    st = np.zeros(6, dtype=np.float32)
    sml = mp.DNAMemorySML()
    sml.Create(a, seed)
    lib.mcu_sml_last_stats(st.ctypes.data)
    assert st[2] > 0 and int(st[3]) >= 1 and int(st[5]) == len(a) - mp.getSeedLength(seed) + 1
    s.close()
