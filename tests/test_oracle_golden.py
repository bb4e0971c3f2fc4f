"""CPU: the C restatement (oracle/mauve_oracle.c) against the golden vectors minted from the reference's own code."""
import hashlib

import numpy as np
import pytest

import _golden


def test_seed_tables(orc):
    t = _golden.seeds()
    for k, v in t["get_seed"].items():
        w, r = (int(x) for x in k.split(","))
        assert orc.get_seed(w, r) == v, k
    for k, v in t["seed_length"].items():
        assert orc.seed_length(int(k)) == v
    for k, v in t["seed_weight"].items():
        assert orc.seed_weight(int(k)) == v
    for k, v in t["default_weight"].items():
        assert orc.default_seed_weight(int(k)) == v, k


def test_sml_small(orc):
    z = _golden.npz("sml_small.npz")
    for name, w, r, seed in _golden.cases(z):
        seq = z["seq_" + name].tobytes()
        pos, mer = orc.sml_build(seq, seed)
        key = "%s_w%d_r%d" % (name, w, r)
        assert np.array_equal(mer, z["mer_" + key]), key
        assert np.array_equal(pos, z["pos_" + key]), key  # oracle emits ties position-ascending = canonical form


def test_mums_small(orc):
    z = _golden.npz("mums_small.npz")
    for i, w, r, rule, seed, coll, cnt in _golden.cases(z):
        rows, stats = orc.find_mums(z["a%d" % i].tobytes(), z["b%d" % i].tobytes(), seed, rule)
        assert np.array_equal(rows, z["rows%d" % i]), i
        assert int(stats[1]) == cnt


def test_mums_mds42(orc):
    z = _golden.npz("mums_mds42.npz")
    m = _golden.meta(z)
    g0, g1 = _golden.mds42()
    rows, stats = orc.find_mums(g0, g1, m["w15_r3"]["seed"], 0)
    assert rows.shape[0] == 29403 and int(rows[:, 0].sum()) == 3792460 and int((rows[:, 2] < 0).sum()) == 1515
    assert np.array_equal(rows, z["rows_w15_r3"])
    assert int(stats[3]) == m["w15_r3"]["n"] + m["w15_r3"]["collisions"] == 2744091  # seed pairs = matches + collisions


def test_sml_mds42_digest(orc):
    z = _golden.npz("mums_mds42.npz")
    m = _golden.meta(z)["sml_full_w15_r3"]
    _, g1 = _golden.mds42()
    pos, mer = orc.sml_build(g1, 0x16df6d)
    assert mer.size == m["n"]
    assert hashlib.sha1(mer.tobytes()).hexdigest() == m["sha1_mer"]
    assert hashlib.sha1(pos.tobytes()).hexdigest() == m["sha1_pos_canon"]


def test_nw_small(orc):
    z = _golden.npz("nw_small.npz")
    for i in range(int(z["n"])):
        path, score = orc.nw_align(z["a%d" % i].tobytes(), z["b%d" % i].tobytes())
        assert path == z["p%d" % i].tobytes(), i


def test_hmm_small(orc):
    z = _golden.npz("hmm_small.npz")
    for c in _golden.cases(z):
        i = c[0]
        sym, params = z["sym%d" % i].tobytes(), z["params%d" % i]
        # parameters: same doubles as getAdaptedHoxdMatrixParameters + adaptToPercentIdentity
        assert np.array_equal(orc.hmm_params(c[1], c[2], c[3], c[4]), params)
        pred, post = orc.hmm_run(sym, params)
        ref_post = z["post%d" % i]
        # the restatement performs the reference's bfloat operations (float32 mantissa) in the same order: bit-identical
        assert np.array_equal(post, ref_post), (i, np.max(np.abs(post - ref_post) / np.maximum(ref_post, 1e-300)))
        assert pred == z["pred%d" % i].tobytes()


# ---- the full-size property checkers (tests/_properties.py) are validated here against oracle output ----
@pytest.mark.parametrize("w,r,n", [(19, 3, 250000), (15, 3, 200000), (11, 0, 120000)])
def test_property_checkers_accept_oracle_mums_and_reject_damage(orc, w, r, n):
    import _properties as P
    import mauve_py_b200 as mp
    from mauve_py_b200 import synth
    a, b = synth.small_pair(n, seed=70 + w, snp=0.02, n_inv=3)
    seed = mp.getSeed(w, r)
    L = mp.getSeedLength(seed)
    rows, _ = orc.find_mums(a, b, seed, 0)
    assert rows.shape[0] > 50 and (rows[:, 2] < 0).any()
    assert P.check_mum_rows(a, b, rows, seed, L) == rows.shape[0]
    for damage in ("shorten", "shift", "swap"):
        bad = rows.copy()
        k = rows.shape[0] // 2
        if damage == "shorten":
            bad[k, 0] -= 1       # no longer maximal
        elif damage == "shift":
            bad[k, 1] += 1       # end seeds are no hits (or order breaks)
        else:
            bad[[0, -1]] = bad[[-1, 0]]
        with pytest.raises(AssertionError):
            P.check_mum_rows(a, b, bad, seed, L)


def test_property_checkers_nw_and_mers(orc):
    import _properties as P
    import mauve_py_b200 as mp
    from mauve_py_b200 import synth
    for x, y in synth.dp_pairs(25, 5, 600, seed=9) + [(b"A", b"ACGT"), (b"ACGTT", b"C"), (b"A", b"A"), (b"AC", b"A"), (b"AAAA", b"TTTTGGGG")]:
        path, score = orc.nw_align(x, y)
        assert P.nw_path_score(x, y, path) == score
    a, _ = synth.small_pair(60000, seed=4)
    for w, r in ((15, 3), (21, 0), (7, 0)):
        seed = mp.getSeed(w, r)
        L, wt = mp.getSeedLength(seed), mp.getSeedWeight(seed)
        pos, mer = orc.sml_build(a, seed)
        assert np.array_equal(P.canonical_mers(a, seed, L, wt, pos), mer)


def _path_from_predicates(bits, last, la, lb):
    """nw_traceback_kernel's walk (csrc/dp.cu) over the four predicates per cell: bit0 = best is not M, bit1 = I beats D,
    bit2 = D came from M, bit3 = I came from M; cells of row / column 1 take the kernel's first-row / first-column rule"""
    M, D, I = last
    edge, sc = "M", M
    if D > sc:
        edge, sc = "D", D
    if I > sc:
        edge, sc = "I", I
    pa, pb, out = la, lb, []
    nib = lambda i, j: int(bits[i - 1, j - 1])
    while True:
        out.append(edge)
        if edge == "M":
            if pa >= 2 and pb >= 2:
                nb = nib(pa - 1, pb - 1)
                x = (nb & 1) + (nb & (nb >> 1) & 1)
                nxt = "MDI"[x]
            else:
                nxt = "D" if pa >= 2 else "I"
            pa, pb = pa - 1, pb - 1
        elif edge == "D":
            nxt = "M" if (pb >= 1 and nib(pa, pb) & 4) else "D"
            pa -= 1
        else:
            nxt = "M" if (pa >= 1 and nib(pa, pb) & 8) else "I"
            pb -= 1
        if pa == 0 and pb == 0:
            break
        edge = nxt
        assert not ((edge == "M" and (pa == 0 or pb == 0)) or (edge == "D" and pa == 0) or (edge == "I" and pb == 0))
    return "".join(reversed(out)), sc


def test_integer_recurrence_of_the_dp_kernel_equals_nwsmall_on_every_tiny_region(orc):
    """the integer recurrence csrc/dp.cu is built on (its header: D' = D - 200, I' = I - 200, terminal gaps folded into the first row
    and column, la == 1 special case) and the kernel's traceback walk, restated in plain Python (tests/_properties.py), give the
    oracle's NWSmall path and score for EVERY pair of sequences up to 3 x 3 (7,056 pairs: all first-row / first-column / la == 1 /
    lb == 1 cases) and for random longer ones; the biased form proposed in experiments/README.md gives the same predicates"""
    import itertools
    import _properties as P
    L = b"ACGT"
    todo = [(a, b) for la in (1, 2, 3) for lb in (1, 2, 3) for a in itertools.product(range(4), repeat=la) for b in itertools.product(range(4), repeat=lb)]
    rng = np.random.default_rng(12)
    todo += [(tuple(rng.integers(0, 4, int(rng.integers(1, 40)))), tuple(rng.integers(0, 4, int(rng.integers(1, 40))))) for _ in range(60)]
    for a, b in todo:
        bits, last = P.nw_integer_recurrence(np.array(a), np.array(b), biased=False)
        path, sc = _path_from_predicates(bits, last, len(a), len(b))
        opath, osc = orc.nw_align(bytes(L[i] for i in a), bytes(L[i] for i in b))
        opath = opath.decode() if isinstance(opath, bytes) else "".join(opath)
        assert sc == osc and path == opath, (a, b, path, opath, sc, osc)
        if len(a) * len(b) > 4:
            bits2, last2 = P.nw_integer_recurrence(np.array(a), np.array(b), biased=True)
            assert np.array_equal(bits, bits2) and (last == last2 or min(last) < P.NW_NINF // 2)


def _real_dp_calls():
    import ast
    z = _golden.npz("dp_mds42_calls.npz")
    split = lambda k: z[k].tobytes().split(b"\n")
    return list(zip(split("a"), split("b"), split("path"))), ast.literal_eval(str(z["meta"]))


def test_nw_real_pipeline_calls(orc):
    """gapped-DP calls recorded from the reference binary aligning the MDS42 pair (tests/golden/make_golden_dp.py): all 61,773
    GlobalAlign calls of that run are of the single-sequence ACGT form; on the committed sample the restatement returns the
    reference's path"""
    calls, meta = _real_dp_calls()
    assert meta["calls"] == meta["single_sequence_acgt_calls"] == 61773 and len(calls) >= 1300
    for a, b, p in calls:
        assert orc.nw_align(a, b)[0] == p


def _real_gap_calls():
    import ast
    z = _golden.npz("gaps_mds42_calls.npz")
    s0, s1 = z["seq0"].tobytes().split(b"\n"), z["seq1"].tobytes().split(b"\n")
    bounds = np.concatenate(([0], np.cumsum(z["counts"])))
    rows = [z["rows"][bounds[i]:bounds[i + 1]] for i in range(len(s0))]
    return s0, s1, z["seeds"], rows, ast.literal_eval(str(z["meta"]))


def test_gap_searches_of_the_real_pipeline(orc):
    """the 371 MemHash::FindMatches calls recursive anchoring makes while the reference aligns the MDS42 pair (recorded with a
    link-time tap, tests/golden/make_golden_taps.py): same rows, same order"""
    s0, s1, seeds, rows, meta = _real_gap_calls()
    assert len(s0) == meta["gap_searches"] == 371 and sum(r.shape[0] for r in rows) == meta["matches"]
    for a, b, seed, want in zip(s0, s1, seeds, rows):
        got, _ = orc.find_mums(a, b, int(seed), 1)
        assert np.array_equal(got, want.reshape(-1, 3))


def test_hmm_call_of_the_real_pipeline(orc):
    """the backbone stage makes ONE run() call for the MDS42 pair, on a 3,983,034-column string; its first million columns with
    the prediction of the reference's run(): the restatement's calls are identical"""
    import ast
    z = _golden.npz("hmm_mds42_call.npz")
    meta = ast.literal_eval(str(z["meta"]))
    assert meta["run_calls"] == 1 and meta["columns_per_call"] == [3983034]
    pred, post = orc.hmm_run(z["sym"].tobytes(), z["params"])
    assert pred == z["pred"].tobytes() and pred.count(b"H") == meta["prefix_homologous"]
