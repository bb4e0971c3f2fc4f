"""Deterministic synthetic inputs for the BASELINE.json configurations.

All generators use numpy.random.Generator(PCG64(seed)) and return ASCII `bytes`/uint8 arrays of
upper-case ACGT.  Shared by tests/, bench.py and __graft_entry__.smoke(); nothing here computes
any part of the hot path.
"""
import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
_COMP[ACGT] = np.frombuffer(b"TGCA", dtype=np.uint8)
_TRANSITION = np.zeros(256, dtype=np.uint8)
_TRANSITION[ACGT] = np.frombuffer(b"GTAC", dtype=np.uint8)
_CODE = np.zeros(256, dtype=np.uint8)
_CODE[ACGT] = np.arange(4, dtype=np.uint8)


def rng_for(seed):
    return np.random.Generator(np.random.PCG64(seed))


def random_genome(n, gc, rng):
    half_gc, half_at = gc / 2.0, (1.0 - gc) / 2.0
    u = rng.random(n, dtype=np.float32)
    codes = np.zeros(n, dtype=np.uint8)  # A
    codes[u >= half_at] = 1  # C
    codes[u >= half_at + half_gc] = 2  # G
    codes[u >= half_at + 2 * half_gc] = 3  # T
    return ACGT[codes]


def revcomp(seq):
    return _COMP[seq[::-1]]


def snps(seq, rate, rng, ts_fraction=2.0 / 3.0):
    """point substitutions at `rate` per base; transitions with probability ts_fraction (2:1 ts:tv)"""
    out = seq.copy()
    pos = np.flatnonzero(rng.random(seq.size, dtype=np.float32) < rate)
    if pos.size == 0:
        return out
    old = out[pos]
    ts = rng.random(pos.size) < ts_fraction
    new = np.where(ts, _TRANSITION[old], 0).astype(np.uint8)
    tv = np.flatnonzero(~ts)
    if tv.size:
        # the two transversion targets of a base: the bases of the other ring class
        oc = _CODE[old[tv]]
        pick = rng.integers(0, 2, tv.size)
        purine = (oc == 0) | (oc == 2)
        cand = np.where(purine, np.where(pick == 0, 1, 3), np.where(pick == 0, 0, 2))
        new[tv] = ACGT[cand]
    out[pos] = new
    return out


def codon_recode(seq, n_windows, window, p, rng):
    """in `n_windows` random in-frame windows replace every third base by its transition with probability p"""
    out = seq.copy()
    if seq.size <= window:
        return out
    starts = rng.integers(0, (seq.size - window) // 3, n_windows) * 3
    for s in starts:
        idx = np.arange(s + 2, s + window, 3)
        hit = idx[rng.random(idx.size) < p]
        out[hit] = _TRANSITION[out[hit]]
    return out


def indels(seq, n_events, geom_p, max_len, rng):
    if n_events == 0:
        return seq.copy()
    pos = np.sort(rng.integers(0, seq.size, n_events))
    lens = np.minimum(rng.geometric(geom_p, n_events), max_len)
    is_ins = rng.random(n_events) < 0.5
    parts, cur = [], 0
    for p, l, ins in zip(pos, lens, is_ins):
        if p < cur:
            continue
        parts.append(seq[cur:p])
        if ins:
            parts.append(ACGT[rng.integers(0, 4, l)])
            cur = p
        else:
            cur = min(p + l, seq.size)
    parts.append(seq[cur:])
    return np.concatenate(parts)


def inversions(seq, n, lo, hi, rng):
    out = seq.copy()
    for _ in range(n):
        l = int(rng.integers(lo, hi + 1))
        if l >= out.size:
            continue
        s = int(rng.integers(0, out.size - l))
        out[s:s + l] = revcomp(out[s:s + l])
    return out


def translocations(seq, n, lo, hi, rng):
    """reciprocal translocations: swap two equal-length, non-overlapping segments in place"""
    out = seq.copy()
    for _ in range(n):
        l = int(rng.integers(lo, hi + 1))
        if 2 * l >= out.size:
            continue
        s1 = int(rng.integers(0, out.size - 2 * l))
        s2 = int(rng.integers(s1 + l, out.size - l + 1))
        tmp = out[s1:s1 + l].copy()
        out[s1:s1 + l] = out[s2:s2 + l]
        out[s2:s2 + l] = tmp
    return out


def small_pair(n, seed=1, snp=0.02, n_indels=None, n_inv=1):
    """small divergent pair for parity tests: SNPs, indels, inversions"""
    rng = rng_for(seed)
    a = random_genome(n, 0.5, rng)
    b = snps(a, snp, rng)
    b = indels(b, n // 2000 if n_indels is None else n_indels, 0.2, 30, rng)
    if n_inv and n > 4000:
        b = inversions(b, n_inv, n // 20, n // 8, rng)
    return a.tobytes(), b.tobytes()


def config2_pair(n=5_000_000, seed=20261017):
    """C2: 5 Mbp bacterial pair -- 1 % SNPs (2:1 ts:tv), codon recoding in 4,000 1-kb windows (p = 0.7), 500 indels <= 30 bp"""
    rng = rng_for(seed)
    a = random_genome(n, 0.508, rng)
    b = snps(a, 0.01, rng)
    b = codon_recode(b, max(1, int(4000 * n / 5_000_000)), 1000, 0.7, rng)
    b = indels(b, max(1, int(500 * n / 5_000_000)), 0.2, 30, rng)
    return a, b


def config3_pair(n=100_000_000, seed=20261018):
    """C3: 100 Mbp pair -- 0.9 % SNP + 0.1 % indel events, 200 inversions (10 kb-1 Mb), 100 translocations (10 kb-500 kb)"""
    rng = rng_for(seed)
    scale = n / 100_000_000
    a = random_genome(n, 0.41, rng)
    b = snps(a, 0.009, rng)
    b = indels(b, int(0.001 * n), 0.3, 30, rng)
    hi_inv = max(2000, int(1_000_000 * min(1.0, scale * 4)))
    hi_tr = max(2000, int(500_000 * min(1.0, scale * 4)))
    b = inversions(b, max(1, int(200 * scale)), min(10_000, hi_inv // 2), hi_inv, rng)
    b = translocations(b, max(1, int(100 * scale)), min(10_000, hi_tr // 2), hi_tr, rng)
    return a, b


def dp_pairs(count, lo, hi, seed=20261020, snp=0.05, indel=0.01):
    """C5: region pairs, lenA log-uniform in [lo, hi]; B = A with 5 % SNPs and 1 % indel events (Geom(0.3) lengths)"""
    rng = rng_for(seed)
    la = np.exp(rng.uniform(np.log(lo), np.log(hi), count)).astype(np.int64)
    la = np.clip(la, lo, hi)
    out = []
    for l in la:
        a = random_genome(int(l), 0.5, rng)
        b = snps(a, snp, rng)
        b = indels(b, int(rng.binomial(int(l), indel)), 0.3, 30, rng)
        if b.size == 0:
            b = a[:1].copy()
        out.append((a.tobytes(), b.tobytes()))
    return out


def dp_arrays(pairs):
    """pairs -> (a, a_off, b, b_off) arrays for nw_batch_arrays"""
    n = len(pairs)
    a_off = np.zeros(n + 1, dtype=np.uint64)
    b_off = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum([len(p[0]) for p in pairs], out=a_off[1:])
    np.cumsum([len(p[1]) for p in pairs], out=b_off[1:])
    a = np.frombuffer(b"".join(p[0] for p in pairs), dtype=np.uint8)
    b = np.frombuffer(b"".join(p[1] for p in pairs), dtype=np.uint8)
    return a, a_off, b, b_off


def hmm_string(n, seed=1, block=400):
    """symbol string '1'..'8' alternating homologous-looking and unrelated-looking blocks"""
    rng = rng_for(seed)
    p_h = np.array([0.33, 0.28, 0.07, 0.16, 0.035, 0.027, 0.004, 0.094])
    p_u = np.array([0.09, 0.09, 0.175, 0.175, 0.09, 0.09, 0.05, 0.24])
    p_h, p_u = p_h / p_h.sum(), p_u / p_u.sum()
    out = np.empty(n, dtype=np.uint8)
    i, homolog = 0, True
    while i < n:
        l = int(rng.integers(block // 4, block * 2))
        k = min(l, n - i)
        out[i:i + k] = rng.choice(8, k, p=p_h if homolog else p_u) + ord("1")
        i += k
        homolog = not homolog
    return out.tobytes()


def colliding_diagonals_pair(n=120_000, n_copies=150, seed=5, snp=0.01, table=40000):
    """Pair built to make hash buckets order dependent (csrc/replay.cu): genome 1 = genome 0 with SNPs (main diagonal,
    offset 0) plus, in a random tail, original 41-bp windows around SNPs placed at offsets that are multiples of the
    reference's hash-table size, so their matches share bucket 0 with the main diagonal and start inside its spans."""
    rng = rng_for(seed)
    a = random_genome(n, 0.5, rng)
    b = snps(a, snp, rng)
    diff = np.flatnonzero(a != b)
    diff = diff[(diff > 100) & (diff < n - 100)]
    pick = rng.choice(diff, min(n_copies, diff.size), replace=False)
    tail_len = table * 4
    tail = random_genome(tail_len, 0.5, rng)
    base = ((n + table - 1) // table) * table  # first multiple of the table size at or after the end of genome 1
    for j, x in enumerate(np.sort(pick)):
        k = int(rng.integers(0, 3))
        q = (x - 20) + base + k * table - n  # index into the tail: genome-1 position = q + n, diagonal = base + k*table
        if 0 <= q and q + 41 <= tail_len:
            tail[q:q + 41] = a[x - 20:x + 21]
    return a.tobytes(), np.concatenate([b, tail]).tobytes()


def repeat_rich_pair(n=300_000, unit=61, copies=3000, seed=11, snp=0.01):
    """Pair with a long tandem repeat (every mer inside the unit occurs `copies` times) and dispersed copies of a
    second element: exercises the overflow / spill paths of the bucketed enumeration (csrc/bucket.cu) and
    MER_REPEAT_LIMIT-sized runs (LM/MatchFinder.cpp:166)."""
    rng = rng_for(seed)
    a = random_genome(n, 0.5, rng)
    u = random_genome(unit, 0.5, rng)
    tandem = np.tile(u, copies)
    k = n // 3
    a = np.concatenate([a[:k], tandem, a[k:]])
    elem = random_genome(400, 0.5, rng)
    for p in rng.integers(0, a.size - 400, 700):
        a[p:p + 400] = elem
    b = snps(a, snp, rng)
    return a.tobytes(), b.tobytes()


def alignment_window(ncol, seed=1, n_rows=2, snp=0.1, gap_rate=0.01, gap_mean=12, both_gap=0.002, wildcards=0.002, term_gaps=True,
                     diverged_blocks=True):
    """an alignment window as MUSCLE's anchor-column search sees it: uint8[n_rows, ncol] of ACGT with '-' runs, a few
    columns where all rows have a gap, IUPAC wildcards and lower case, stretches that are well conserved and stretches that are not"""
    rng = rng_for(seed)
    base = random_genome(ncol, 0.5, rng)
    rows = np.empty((n_rows, ncol), dtype=np.uint8)
    for r in range(n_rows):
        row = base.copy()
        rate = np.full(ncol, snp if r else snp / 4, dtype=np.float32)
        if diverged_blocks:
            pos = 0
            while pos < ncol:
                ln = int(rng.integers(30, 400))
                if rng.random() < 0.3:
                    rate[pos:pos + ln] = 0.6
                pos += ln
        hit = rng.random(ncol, dtype=np.float32) < rate
        row[hit] = ACGT[rng.integers(0, 4, int(hit.sum()))]
        n_gaps = int(rng.poisson(gap_rate * ncol))
        for _ in range(n_gaps):
            ln = int(rng.geometric(1.0 / gap_mean))
            at = int(rng.integers(0, max(ncol - 1, 1)))
            row[at:at + ln] = ord("-")
        if wildcards > 0:
            w = rng.random(ncol, dtype=np.float32) < wildcards
            keep = row != ord("-")
            row[w & keep] = np.frombuffer(b"NNNRYKMSWn", dtype=np.uint8)[rng.integers(0, 10, int((w & keep).sum()))]
            low = (rng.random(ncol, dtype=np.float32) < wildcards) & keep & ~w
            row[low] = row[low] | 0x20
        if term_gaps and ncol > 8 and rng.random() < 0.5:
            k = int(rng.integers(1, max(ncol // 8, 2)))
            if rng.random() < 0.5:
                row[:k] = ord("-")
            else:
                row[ncol - k:] = ord("-")
        rows[r] = row
    if both_gap > 0 and ncol:
        g = rng.random(ncol, dtype=np.float32) < both_gap
        for s in np.nonzero(g)[0]:
            rows[:, s:s + int(rng.integers(1, 6))] = ord("-")
    return rows
