"""buildIndex -- the reference's Python entry point (mauve/buildindex.py:90-138) with the anchoring stage on the GPU.

    idx_lut = mauve_py_b200.buildIndex(genome_fp, ref_genome_fp)      # same signature, same int32 LUT

The reference runs `progressiveMauveStatic a.fa b.fa --output x.xmfa`, parses the XMFA and builds, for every base of genome A,
the 0-based index of the identical aligned base of genome B (or -1), then heals the table (fixZeroIdx, fillGaps, smoothEdges of
mauve/indexutils.pyx:27-108).  Here the initial seed + match + extend pass -- `PairwiseMatchFinder::FindMatches` on two
`DNAFileSML`s, MA/progressiveMauve.cpp:446-503 -- runs on the device (mcu_find_mums) and the UNMODIFIED binary is started with
`--match-input` (MA/progressiveMauve.cpp:472-491 -> ReadList, LM/MatchList.h:526-614) and finds the two `<fasta>.sslist` sorted mer
lists already written from the device's lists, so it skips its own SML construction and match finding and continues with what
it is given; everything after that (LCBs, recursive anchoring, gapped alignment, backbone, XMFA writer)
is the reference's code.  The LUT is bit-identical to `mauve.buildIndex` (tests/golden/mds42_lut.npz was minted by the reference's
own buildIndex, tests/golden/make_golden_lut.py).

As in the reference, the binary is looked up in $MAUVE_DIR (`progressiveMauveStatic`, or `progressiveMauve`); there is no fallback
for the device part: without a GPU the call raises.
"""
import os
import shutil
import subprocess
import tempfile

import numpy as np

from . import libmems

IDX_ARRAY_DTYPE = np.int32


def _binary():
    d = os.environ.get("MAUVE_DIR")
    if not d:
        raise IOError("MAUVE_DIR is not set: it must name the directory holding progressiveMauveStatic (as for the reference package)")
    for name in ("progressiveMauveStatic", "progressiveMauve"):
        fp = os.path.join(d, name)
        if os.path.isfile(fp) and os.access(fp, os.X_OK):
            return fp
    raise IOError("no progressiveMauveStatic in MAUVE_DIR=%s" % d)


def getSeqFromFile(fasta_fp):
    """sequence of the single record of a FASTA file (what libnano.fileio.getSeqFromFile gives buildIndex)"""
    with open(fasta_fp, "rb") as f:
        lines = f.read().split(b"\n")
    if sum(1 for l in lines if l.startswith(b">")) != 1:
        raise ValueError("%s: buildIndex expects one FASTA record per genome" % fasta_fp)
    return b"".join(l.strip() for l in lines if l and not l.startswith(b">")).decode("ascii")


def parseXMFA(xmfa_fp):
    """[[(seq_num, start_idx, end_idx, strand, seq), ...] per '='-terminated block], the fields mauve/xmfa.py:38-70 extracts"""
    groups, cur, entry = [], [], None
    with open(xmfa_fp) as f:
        for line in f:
            if line.startswith("#"):
                continue
            if line.startswith("="):
                if entry is not None:
                    cur.append(entry)
                if cur:
                    groups.append([(n, s, e, st, "".join(parts)) for n, s, e, st, parts in cur])
                cur, entry = [], None
            elif line.startswith(">"):
                if entry is not None:
                    cur.append(entry)
                head = line[1:].split()
                num, rng = head[0].split(":")
                s, e = rng.split("-")
                entry = (int(num), int(s), int(e), head[1], [])
            elif entry is not None:
                entry[4].append(line.strip())
    return groups


def lut_from_alignment(groups, genome_length):
    """the per-column walk of buildindex.py:112-130, including its conventions: the strand is ignored, and a base is entered at
    the index reached AFTER counting it (so block position p lands in slot p, one past its 0-based index)"""
    lut = np.full(genome_length, -1, dtype=IDX_ARRAY_DTYPE)
    for block in groups:
        a = [x for x in block if x[0] == 1]
        b = [x for x in block if x[0] == 2]
        if not a or not b:
            continue
        sa = np.frombuffer(a[0][4].encode("ascii"), dtype=np.uint8)
        sb = np.frombuffer(b[0][4].encode("ascii"), dtype=np.uint8)
        n = min(sa.size, sb.size)
        sa, sb = sa[:n], sb[:n]
        gi = (a[0][1] - 1) + np.cumsum(sa != ord("-"))
        ri = (b[0][1] - 1) + np.cumsum(sb != ord("-"))
        m = (sa == sb) & (gi < genome_length)
        lut[gi[m]] = ri[m].astype(IDX_ARRAY_DTYPE)
    return lut


def fixZeroIdx(idx_lut, genome, ref_genome):
    """mauve/indexutils.pyx:27-43"""
    if idx_lut[0] != -1:
        return
    n = idx_lut.shape[0]
    nxt = 1
    while nxt < n and (idx_lut[nxt] == -1 or nxt < 10):
        nxt += 1
    if nxt >= n or idx_lut[nxt] == -1:
        return
    ref_idx = int(idx_lut[nxt]) - nxt
    if ref_idx > -1 and genome[0] == ref_genome[ref_idx]:
        idx_lut[0] = ref_idx


def _heal(idx_lut, radius, candidates):
    """shared body of fillGaps / smoothEdges: from a candidate position look up to `radius` entries ahead for a mapped entry that is
    as far away in the other genome as in this one, and make the stretch in between linear"""
    n = idx_lut.shape[0]
    for idx in candidates:
        lower = idx - 1
        lm = int(idx_lut[lower])
        if lm == -1:
            continue
        cur = int(idx_lut[idx])
        is_candidate = (cur == -1) if radius[1] else (cur != lm + 1)  # re-checked: an earlier fill may have covered this entry
        if not is_candidate:
            continue
        hi = min(n, idx + radius[0] + 1)
        if hi <= idx + 1:
            continue
        win = idx_lut[idx + 1:hi].astype(np.int64)
        ok = np.flatnonzero(win - lm == np.arange(idx + 1 - lower, hi - lower))
        if ok.size:
            up = idx + 1 + int(ok[0])
            idx_lut[lower:up + 1] = np.arange(lm, int(idx_lut[up]) + 1, dtype=IDX_ARRAY_DTYPE)


def fillGaps(idx_lut, max_gap_width=300):
    """mauve/indexutils.pyx:46-75.  Only the first entry of a run of -1 can start a fill, fills only touch later entries and never
    create a -1, so visiting the run starts in ascending order and re-checking each one is the reference's full scan."""
    v = idx_lut
    starts = np.flatnonzero((v[1:] == -1) & (v[:-1] != -1)) + 1
    _heal(idx_lut, (max_gap_width, True), starts.tolist())


def smoothEdges(idx_lut, smoothing_radius=20):
    """mauve/indexutils.pyx:78-108.  Same argument: an entry that follows its predecessor by one is never touched, a smoothing only
    rewrites later entries into a linear stretch (creating no new edge beyond its end, whose value is kept)."""
    v = idx_lut.astype(np.int64)
    edges = np.flatnonzero((v[1:] != v[:-1] + 1) & (v[:-1] != -1)) + 1
    _heal(idx_lut, (smoothing_radius, False), edges.tolist())


def lut_from_xmfa(xmfa_fp, genome_seq, ref_genome_seq, fill_gaps=True, max_gap_width=300, smooth_edges=True, smoothing_radius=20):
    """buildindex.py:105-138 from the parsed XMFA on"""
    lut = lut_from_alignment(parseXMFA(xmfa_fp), len(genome_seq))
    if fill_gaps:
        fixZeroIdx(lut, genome_seq, ref_genome_seq)
        fillGaps(lut, max_gap_width=max_gap_width)
    if smooth_edges:
        smoothEdges(lut, smoothing_radius=smoothing_radius)
    return lut


def runMauve(fasta_files, flags):
    """runMauve of buildindex.py:56-80: same command line, scratch directories and .sslist clean-up"""
    abs_paths = [os.path.abspath(fp) for fp in fasta_files]
    d1, d2 = tempfile.mkdtemp(), tempfile.mkdtemp()
    try:
        flags = dict(flags)
        flags.update({"--scratch-path-1": d1, "--scratch-path-2": d2})
        args = [_binary()] + abs_paths + [str(el) for pair in flags.items() for el in pair]
        proc = subprocess.Popen(args, stdout=subprocess.PIPE, stdin=subprocess.PIPE, stderr=subprocess.PIPE, cwd=d1)
        out, err = proc.communicate()
        if proc.returncode != 0:
            raise RuntimeError("progressiveMauve failed (%d): %s" % (proc.returncode, err.decode(errors="replace")[-400:]))
    finally:
        for d in (d1, d2):
            shutil.rmtree(d, ignore_errors=True)
        for fp in abs_paths:
            try:
                os.remove(fp + ".sslist")
            except OSError:
                pass


def buildIndex(genome_fp, ref_genome_fp, genome_seq=None, ref_genome_seq=None, fill_gaps=True, max_gap_width=300, smooth_edges=True,
               smoothing_radius=20):
    """Drop-in for mauve.buildIndex (buildindex.py:90-138): index mapping from the genome in `genome_fp` to the one in `ref_genome_fp`.
    Initial anchors and sorted mer lists come from the device; when a mer occurs more than MER_REPEAT_LIMIT times (stats[3] of
    mcu_find_mums) only the lists do and the binary finds its anchors itself (see below)."""
    genome_fp, ref_genome_fp = os.path.abspath(genome_fp), os.path.abspath(ref_genome_fp)
    genome_seq = genome_seq or getSeqFromFile(genome_fp)
    ref_genome_seq = ref_genome_seq or getSeqFromFile(ref_genome_fp)
    # initial anchors on the device: default seed weight from the average length, coding pattern (LM/MatchList.h:265-280)
    weight = libmems.getDefaultSeedWeight((len(genome_seq) + len(ref_genome_seq)) // 2)
    seed = libmems.getSeed(weight, libmems.CODING_SEED)
    seqs = (genome_seq.encode("ascii"), ref_genome_seq.encode("ascii"))
    rows, stats = libmems.find_mums(seqs[0], seqs[1], seed)
    # stats[3] != 0: some mer has more than MER_REPEAT_LIMIT (1000) copies.  There the reference's own list carries a few rows whose
    # existence depends on std::sort's tie order (include/mauve_cuda.h); the device list leaves them out.  To keep
    # the LUT the reference binary's in that case too, the anchors are then NOT handed over: the unmodified binary runs its own
    # PairwiseMatchFinder on the sorted mer lists written below (the lists themselves are identical either way).
    repeat_limit_hit = int(stats[3]) != 0
    # the sorted mer lists the binary would build next to the FASTA files (DNAFileSML::Create, LM/FileSML.cpp:401-459): written from
    # the device's lists, MatchList::LoadSMLs loads them instead (LM/MatchList.h:296-330); runMauve removes them afterwards
    for fp, seq in zip((genome_fp, ref_genome_fp), seqs):
        sml = libmems.DNAMemorySML()
        sml.Create(seq, seed)
        sml.WriteFile(fp + ".sslist")
    work = tempfile.mkdtemp()
    try:
        flags = {"--output": os.path.join(work, "mauveout.xmfa")}
        if rows.shape[0] and not repeat_limit_hit:  # WriteList prints nothing for an empty list (LM/MatchList.h:619-620): let the binary search itself then
            mums_fp = os.path.join(work, "anchors.mums")
            with open(mums_fp, "w") as f:
                libmems.WriteList(rows, f, (genome_fp, ref_genome_fp), (len(genome_seq), len(ref_genome_seq)))
            flags["--match-input"] = mums_fp
        runMauve([genome_fp, ref_genome_fp], flags)
        return lut_from_xmfa(flags["--output"], genome_seq, ref_genome_seq, fill_gaps, max_gap_width, smooth_edges, smoothing_radius)
    finally:
        shutil.rmtree(work, ignore_errors=True)
