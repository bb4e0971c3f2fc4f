"""Size-independent properties of the path's outputs (pure numpy), used where the oracle is too slow: BASELINE.json's full sizes.

Every checker here is itself validated on the CPU against oracle output (tests/test_oracle_golden.py::test_property_checkers_*),
so that a failure at full size points at the device result and not at the checker.  Semantics: SURVEY.md Appendix B.
"""
import numpy as np

_CODE = np.zeros(256, dtype=np.uint8)
_CODE[np.frombuffer(b"ACGT", dtype=np.uint8)] = np.arange(4, dtype=np.uint8)
_CODE[np.frombuffer(b"acgt", dtype=np.uint8)] = np.arange(4, dtype=np.uint8)
for _c, _v in ((b"BY", 1), (b"SK", 2)):  # SortedMerList::BasicDNATable, LM/SortedMerList.cpp:29-47
    _CODE[np.frombuffer(_c, dtype=np.uint8)] = _v
    _CODE[np.frombuffer(_c.lower(), dtype=np.uint8)] = _v


def codes(seq):
    return _CODE[np.frombuffer(seq, dtype=np.uint8) if isinstance(seq, (bytes, bytearray)) else np.asarray(seq, dtype=np.uint8)]


def seed_care_positions(seed, length):
    """offsets (from the seed's first base) of the pattern's 1-positions; pattern MSB = first base (LM/SortedMerList.cpp:726-762)"""
    return np.array([j for j in range(length) if (seed >> (length - 1 - j)) & 1], dtype=np.int64)


def _hits(ca, cb, idx, L, rev, d, ts):
    """hit(t) of SURVEY.md B.2 for the offsets `ts` on diagonal d (all within range)"""
    ts = np.asarray(ts, dtype=np.int64)
    A = ca[ts[:, None] + idx[None, :]]
    if not rev:
        B = cb[(ts + d)[:, None] + idx[None, :]]
        return np.all(A == B, axis=1)
    other = d - ts
    B = 3 - cb[other[:, None] + (L - 1 - idx)[None, :]]
    same = np.all(A == B, axis=1)
    # the reference's parity test (LM/MatchFinder.h:281-303): a seed that is its own reverse complement never matches in reverse
    selfrc = np.all(A == (3 - A[:, ::-1]), axis=1)
    return same & ~selfrc


def check_mum_rows(a, b, rows, seed, L, sample=None, rng=None):
    """Each sampled row is a maximal chain of spaced-seed hits with gaps <= L (its first and last seed are hits, no hit within L
    beyond either end), inside both sequences; the list is in GetMatchList order: ascending (offset mod 40000, start0).
    Returns the number of rows examined; raises AssertionError with the offending row."""
    ca, cb = codes(a), codes(b)
    n0, n1 = ca.size, cb.size
    idx = seed_care_positions(seed, L)
    rows = np.asarray(rows, dtype=np.int64).reshape(-1, 3)
    ln, s0, s1 = rows[:, 0], rows[:, 1], rows[:, 2]
    assert np.all(ln >= L) and np.all(s0 >= 1) and np.all(s0 - 1 + ln <= n0), "genome-0 span out of range"
    assert np.all(np.abs(s1) >= 1) and np.all(np.abs(s1) - 1 + ln <= n1), "genome-1 span out of range"
    off = s1 - s0 - np.where(s1 < 0, ln, 0)
    key = (off % 40000) * (1 << 34) + s0
    assert np.all(key[1:] >= key[:-1]), "rows are not in (offset mod 40000, start0) order"
    pick = np.arange(rows.shape[0])
    if sample is not None and sample < rows.shape[0]:
        pick = np.sort((rng or np.random.default_rng(0)).choice(rows.shape[0], sample, replace=False))
    for r in pick:
        length, lo, rev = int(ln[r]), int(s0[r]) - 1, bool(s1[r] < 0)
        hi = lo + length - L
        d = hi + (-int(s1[r]) - 1) if rev else (int(s1[r]) - 1) - lo
        # offsets at which both seed windows lie inside the sequences
        if rev:
            tmin, tmax = max(0, d - (n1 - L)), min(n0 - L, d)
        else:
            tmin, tmax = max(0, -d), min(n0 - L, n1 - L - d)
        assert tmin <= lo <= hi <= tmax, ("match leaves the valid range", rows[r])
        t0, t1 = max(tmin, lo - L), min(tmax, hi + L)
        ts = np.arange(t0, t1 + 1)
        h = _hits(ca, cb, idx, L, rev, d, ts)
        inside = h[lo - t0:hi - t0 + 1]
        assert inside[0] and inside[-1], ("end seeds are not hits", rows[r])
        assert not h[:lo - t0].any() and not h[hi - t0 + 1:].any(), ("not maximal: a hit within L of an end", rows[r])
        pos = np.flatnonzero(inside)
        assert pos.size == 1 or int(np.max(np.diff(pos))) <= L, ("chain broken: gap > L between hits", rows[r])
    return int(pick.size)


NUC_SP = np.array([[151, -54, 29, -63], [-54, 160, -65, 29], [29, -65, 160, -54], [-63, 29, -54, 151]], dtype=np.int64)  # MU/nucmx.cpp:8-25 (+60)


def nw_path_score(a, b, edges):
    """Score of a global alignment path under the restated NWSmall model (substitution NUC_SP; a gap run costs 400, 200 when it
    touches either end of the path: terminal open or close is free, MU/termgaps.cpp:19-33).  Also checks that the path consumes
    both sequences.  edges: bytes of 'M','D','I'."""
    e = np.frombuffer(edges, dtype=np.uint8)
    ca, cb = codes(a), codes(b)
    useA = (e == ord("M")) | (e == ord("D"))
    useB = (e == ord("M")) | (e == ord("I"))
    assert int(useA.sum()) == ca.size and int(useB.sum()) == cb.size, "path does not consume both sequences"
    ia, ib = np.cumsum(useA) - 1, np.cumsum(useB) - 1
    m = e == ord("M")
    score = int(NUC_SP[ca[ia[m]], cb[ib[m]]].sum())
    change = np.flatnonzero(np.concatenate(([True], e[1:] != e[:-1])))
    ends = np.concatenate((change[1:], [e.size]))
    for s, t in zip(change.tolist(), ends.tolist()):
        if e[s] != ord("M"):
            score -= 200 if (s == 0 or t == e.size) else 400
    if ca.size == 1:
        score -= 200  # a one-letter A profile keeps one gap-close term: M[1][1] = S - 200 (csrc/dp.cu header, MU/nwsmall.cpp:586-592)
    return score


def canonical_mers(seq, seed, L, w, positions):
    """bmer::mer (canonical seed left-aligned in 64 bits | strand bit) of the seeds starting at `positions`"""
    c = codes(seq).astype(np.uint64)
    idx = seed_care_positions(seed, L)
    positions = np.asarray(positions, dtype=np.int64)
    f = np.zeros(positions.size, dtype=np.uint64)
    r = np.zeros(positions.size, dtype=np.uint64)
    for k, j in enumerate(idx.tolist()):
        base = c[positions + j]
        f |= base << np.uint64(2 * (w - 1 - k))
        r |= (np.uint64(3) - base) << np.uint64(2 * k)
    strand = r < f
    canon = np.where(strand, r, f)
    return (canon << np.uint64(64 - 2 * w)) | strand.astype(np.uint64)


def sol_expected(positions, mers, seed_mask, n, L):
    """SeedOccurrenceList frequencies from a sorted mer list, evaluated independently of the device code and of the restatement:
    multiplicities from run lengths, exact integer window sums through a cumulative sum, one double division, one rounding."""
    masked = np.asarray(mers, dtype=np.uint64) & np.uint64(seed_mask)
    raw = np.ones(n, dtype=np.int64)
    if masked.size:
        starts = np.flatnonzero(np.concatenate([[True], masked[1:] != masked[:-1]]))
        lens = np.diff(np.concatenate([starts, [masked.size]]))
        raw[np.asarray(positions)] = np.repeat(lens, lens)
    c = np.concatenate([[0], np.cumsum(np.concatenate([np.ones(L - 1, dtype=np.int64), raw]))])
    want = ((c[L:] - c[:-L]).astype(np.float64) / float(L)).astype(np.float32)
    want[n - 1] = np.float32(raw[n - 1])
    return want


# ---- the integer recurrence csrc/dp.cu is built on (header comment there), cell by cell in plain Python -----------------------------
# biased=False: as the kernel evaluates it today (D' = D - 200, I' = I - 200, M - 400 carried); biased=True: the form of
# experiments/README.md (gap states + 400, M carried).  Returns the four traceback predicates per cell and the final (M, D, I).
NW_NINF = -(1 << 29)
NW_SUB = np.array([[151, -54, 29, -63], [-54, 160, -65, 29], [29, -65, 160, -54], [-63, 29, -54, 151]])  # NUC_SP + 60 (MU/nucmx.cpp:8-25), A C G T


def nw_integer_recurrence(a, b, biased=False):
    la, lb = len(a), len(b)
    bias = 400 if biased else 0
    # row 0 / column 0 as in dp.cu: best[0][0] = 0 (-200 if la == 1), best[i][0] = best[0][j] = -200; no M, D', I' outside the matrix
    best_prev = [(-200 if la == 1 else 0)] + [-200] * lb
    Mrow_prev = [NW_NINF] * (lb + 1)      # the CARRIED M of row i-1: M - 400 today, M in the biased form; NW_NW_NINF outside the matrix in both
    Drow_prev = [NW_NINF] * (lb + 1)      # D state of row i-1 (biased or not)
    bits = np.zeros((la, lb), dtype=np.uint8)
    last = None
    for i in range(la):
        best_row = [-200] + [0] * lb
        Mrow = [NW_NINF] * (lb + 1)
        Drow = [NW_NINF] * (lb + 1)
        I_left, M_left = NW_NINF, NW_NINF
        for j in range(1, lb + 1):
            M = NW_SUB[a[i], b[j - 1]] + best_prev[j - 1]
            upM, upD = Mrow_prev[j], Drow_prev[j]
            D = max(upD, upM)
            I = max(I_left, M_left)
            if biased:
                best = max(max(D, I) - 400, M)
                b0 = best > M
                carried = M
            else:
                best = max(M, max(D, I))
                b0 = max(D, I) > M
                carried = M - 400
            b1, b2, b3 = I > D, upM >= upD, M_left >= I_left
            bits[i, j - 1] = b0 | (b1 << 1) | (b2 << 2) | (b3 << 3)
            Mrow[j], Drow[j], best_row[j] = carried, D, best
            I_left, M_left = I, carried
            last = (M, D + 200 - bias, I + 200 - bias)   # what nw_region stores as the result: (M, D, I) = (M, D' + 200, I' + 200)
        best_prev, Mrow_prev, Drow_prev = best_row, Mrow, Drow
    return bits, last
