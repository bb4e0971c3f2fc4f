"""GPU (-m gpu): the BASELINE.json configurations at their full sizes.

Config 2 (5 Mbp pair) is still within the oracle's reach and is compared row for row.  Config 3 (100 Mbp pair), a 50 Mbp sorted
mer list (config 4's stage at a size the host can check) and a config-5 batch of DP regions are checked through the
size-independent properties of tests/_properties.py, whose checkers are validated against oracle output on the CPU
(tests/test_oracle_golden.py::test_property_checkers_*).  The file sorts last so that the cheap parity tests run first.
"""
import numpy as np
import pytest

import _golden
import _properties as P
from mauve_py_b200 import synth

pytestmark = pytest.mark.gpu


def test_config2_5mbp_pair_bit_exact(mp, orc):
    """synthetic 5 Mbp bacterial pair, seed weight 15 coding pattern: every row, in order, equals the oracle's"""
    a, b = synth.config2_pair()
    seed = mp.getSeed(mp.getDefaultSeedWeight((a.size + b.size) // 2), mp.CODING_SEED)
    assert seed == 0x16DF6D
    rows, stats = mp.libmems.find_mums(a.tobytes(), b.tobytes(), seed)
    orows, ostats = orc.find_mums(a.tobytes(), b.tobytes(), seed, 0)
    assert rows.shape[0] > 10000 and np.array_equal(rows, orows)
    assert int(stats[2]) == int(ostats[0])  # MemCollisionCount


def test_config3_100mbp_pair_properties(mp):
    """synthetic 100 Mbp pair (inversions, translocations, 1 % divergence), default weight 19 -> solid seed"""
    a, b = synth.config3_pair()
    ab, bb = a.tobytes(), b.tobytes()
    seed = mp.getSeed(mp.getDefaultSeedWeight((a.size + b.size) // 2), mp.CODING_SEED)
    L = mp.getSeedLength(seed)
    assert L == 19
    s = mp.AnchorSession()
    s.upload(ab, bb)
    n = s.run(seed)
    rows = s.download().copy()
    assert n == rows.shape[0] > 500000
    # the WHOLE list equals the list the reference's own code returns for this pair (tests/golden/make_golden_config3.py ran
    # oracle/_ref = unmodified MatchFinder / MemHash sources on the full 100 Mbp pair: 414 s on one core)
    import hashlib
    import json
    import os
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "config3_rows.json")))
    assert n == g["reference"]["rows"] == 827941
    assert hashlib.sha1(np.ascontiguousarray(rows, dtype=np.int64).tobytes()).hexdigest() == g["reference"]["sha1"]
    assert int(s.stats[3]) == 0 and g["oracle_equals_reference"]       # no mer beyond MER_REPEAT_LIMIT: the documented divergence cannot occur here
    assert int(s.stats[0]) - int(s.stats[1]) == int(s.stats[2])          # seed pairs = matches + collisions
    # every sampled row is a maximal run of seed hits inside both genomes; the whole list is in GetMatchList order
    assert P.check_mum_rows(ab, bb, rows, seed, L, sample=4000, rng=np.random.default_rng(1)) == 4000
    assert (rows[:, 2] < 0).sum() > 1000                                 # the inversions are found on the reverse strand
    # matches are pairwise distinct here (no order-dependent duplicates) and cover most of the genome once
    assert np.unique(rows, axis=0).shape[0] == rows.shape[0] or int(s.stats[7]) > 0
    assert 0.5 * a.size < int(rows[:, 0].sum()) < 1.2 * a.size
    # idempotent
    assert s.run(seed) == n and np.array_equal(s.download(), rows)
    # a sharded run (each "rank" owning half of the seeds, bitmaps combined) emits every match exactly once
    import torch
    from _devwords import DeviceWords as _DeviceWords
    t = mp.AnchorSession()
    t.upload(ab, bb)
    s.enumerate(seed, 0, 2)
    t.enumerate(seed, 1, 2)
    vs, vt = (torch.as_tensor(_DeviceWords(*x.uniq_bitmap()), device="cuda") for x in (s, t))
    total = vs + vt
    vs.copy_(total)
    vt.copy_(total)
    torch.cuda.synchronize()
    n0, n1 = s.finish(uniq_is_global=True), t.finish(uniq_is_global=True)
    assert n0 + n1 == n
    merged_n, _ = s.merge(np.concatenate([s.download().copy(), t.download().copy()], axis=0))
    assert merged_n == n and np.array_equal(s.download(), rows)
    s.close()
    t.close()


@pytest.mark.parametrize("w,r", [(21, 0), (11, 0)])
def test_sorted_mer_list_50mbp_properties(mp, w, r):
    """SML build at 50 Mbp: mers non-decreasing, positions a permutation, ties position-ascending, sampled mers recomputed"""
    g = synth.random_genome(50_000_000, 0.5, synth.rng_for(50 + w)).tobytes()
    seed = mp.getSeed(w, r)
    L, wt = mp.getSeedLength(seed), mp.getSeedWeight(seed)
    sml = mp.DNAMemorySML()
    sml.Create(g, seed)
    n = len(g) - L + 1
    assert sml.SMLLength() == n
    mers, pos = sml.mers(), sml.positions()
    assert mers.shape[0] == n and pos.shape[0] == n
    assert np.all(mers[1:] >= mers[:-1])
    assert np.array_equal(np.sort(pos), np.arange(n, dtype=pos.dtype))
    tie = mers[1:] == mers[:-1]
    assert np.all(pos[1:][tie] > pos[:-1][tie])
    pick = np.random.default_rng(w).integers(0, n, 200_000)
    assert np.array_equal(P.canonical_mers(g, seed, L, wt, pos[pick]), mers[pick])


def test_config5_dp_batch_properties(mp):
    """a batch of config-5 regions (100 bp - 10 kbp): every path consumes both sequences and its score, recomputed from the path
    under the NWSmall model, is the score the device reports; a sample is compared with the oracle in test_nw_vs_oracle"""
    pairs = synth.dp_pairs(1500, 100, 10000, seed=20261021)
    res = mp.libmems.nw_batch_arrays(*synth.dp_arrays(pairs))
    assert float(res["stats"][0]) == float(sum(len(x) * len(y) for x, y in pairs))
    path, off, plen, score = res["path"], res["path_off"], res["path_len"], res["score"]
    for k, (x, y) in enumerate(pairs):
        edges = path[int(off[k]):int(off[k]) + int(plen[k])].tobytes()
        assert P.nw_path_score(x, y, edges) == int(score[k]), k


def test_nw_real_pipeline_calls(mp):
    """the gapped-DP calls the reference binary really makes while aligning the MDS42 pair (recorded with a link-time tap on
    muscle::GlobalAlign, tests/golden/make_golden_dp.py): one batch, every path equal to the one NWSmall + BitTraceBack returned"""
    import ast
    z = _golden.npz("dp_mds42_calls.npz")
    a, b, want = (z[k].tobytes().split(b"\n") for k in ("a", "b", "path"))
    meta = ast.literal_eval(str(z["meta"]))
    assert meta["calls"] == meta["single_sequence_acgt_calls"] and len(a) >= 1300
    got = mp.GlobalAlignBatch(list(zip(a, b)))
    for g, w in zip(got, want):
        assert g.edges == w


def test_gap_searches_of_the_real_pipeline(mp):
    """the 371 gap searches of the reference's MDS42 run in ONE mcu_find_mums_batch call, with the seeds the reference chose: every
    gap gets the rows MemHash returned (fixture: tests/golden/make_golden_taps.py)"""
    z = _golden.npz("gaps_mds42_calls.npz")
    s0, s1 = z["seq0"].tobytes().split(b"\n"), z["seq1"].tobytes().split(b"\n")
    bounds = np.concatenate(([0], np.cumsum(z["counts"])))
    res, stats = mp.libmems.find_mums_batch(list(zip(s0, s1)), seeds=[int(x) for x in z["seeds"]])
    assert len(res) == 371 and int(stats[1]) == int(z["counts"].sum())
    for i, rows in enumerate(res):
        assert np.array_equal(rows, z["rows"][bounds[i]:bounds[i + 1]].reshape(-1, 3)), i
    # and the seeds are the ones the adapter derives: getSeed(getDefaultSeedWeight((len0 + len1) / 2), 0)
    for a, b, seed in zip(s0, s1, z["seeds"]):
        assert mp.getSeed(mp.getDefaultSeedWeight((len(a) + len(b)) // 2), 0) == int(seed)


def test_hmm_call_of_the_real_pipeline(mp):
    """first million columns of the one genome-sized column string the reference's backbone stage scores for the MDS42 pair:
    H/N calls identical to run()'s"""
    z = _golden.npz("hmm_mds42_call.npz")
    pred = mp.run(z["sym"].tobytes(), z["params"])
    assert pred == z["pred"].tobytes()
