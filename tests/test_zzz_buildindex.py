"""The outermost parity check: the int32 LUT of `mauve.buildIndex(mds42_recoded.fa, mds42_full.fa)`.

tests/golden/mds42_lut.npz was minted by the reference's OWN buildIndex (its Python, its Cython helpers, its binary compiled in
place: tests/golden/make_golden_lut.py).  CPU: the host-side mirror of the Python layer (mauve_py_b200/buildindex.py: XMFA parser,
column walk, fixZeroIdx / fillGaps / smoothEdges) reproduces that LUT from the reference binary's XMFA.  GPU: the drop-in
`mauve_py_b200.buildIndex`, whose initial anchors come from the device and enter the UNMODIFIED binary through --match-input,
returns the same LUT."""
import ast
import gzip
import hashlib
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
BINARY = os.path.join(REF_DIR, "progressiveMauve")

needs_bin = pytest.mark.skipif(not os.path.exists(BINARY), reason="oracle/_ref/progressiveMauve not built (needs /root/reference at build time)")


def _golden():
    z = np.load(os.path.join(GOLDEN, "mds42_lut.npz"))
    lut = np.cumsum(z["lut_diff"].astype(np.int64)).astype(np.int32)
    return lut, ast.literal_eval(str(z["meta"]))


def _fastas(tmp_path):
    out = []
    for name in ("mds42_recoded", "mds42_full"):
        p = os.path.join(str(tmp_path), name + ".fa")
        with gzip.open(os.path.join(GOLDEN, name + ".fa.gz"), "rb") as f, open(p, "wb") as g:
            shutil.copyfileobj(f, g)
        out.append(p)
    return out


def test_golden_lut_is_self_consistent():
    lut, meta = _golden()
    assert lut.size == meta["length"] == 3981477 and int((lut >= 0).sum()) == meta["mapped"]
    assert hashlib.sha1(lut.tobytes()).hexdigest() == meta["lut_sha1"]


def test_cli_match_list_equals_the_golden_list():
    """the list `progressiveMauve --mums` wrote during the LUT run (sha1 in the LUT fixture) is the 29,403-row golden list the oracle
    and the CUDA path are held to (tests/golden/mums_mds42.npz, minted through the library-level driver)"""
    import _golden as G
    _, meta = _golden()
    rows = G.npz("mums_mds42.npz")["rows_w15_r3"]
    text = b"\n".join(b"\t".join(str(int(v)).encode() for v in r) for r in rows)
    assert rows.shape[0] == meta["mums_rows"] == 29403 and hashlib.sha1(text).hexdigest() == meta["mums_sha1"]


def test_lut_healing_on_crafted_tables():
    """fillGaps / smoothEdges against a literal transcription of the reference loops (mauve/indexutils.pyx:46-108)"""
    from mauve_py_b200 import buildindex as B

    def ref_fill(v, width):
        v = v.copy()
        for idx in range(1, len(v)):
            if v[idx] == -1 and v[idx - 1] != -1:
                for up in range(idx + 1, len(v)):
                    if up - idx > width:
                        break
                    if v[up] - v[idx - 1] == up - (idx - 1):
                        v[idx - 1:up + 1] = np.arange(v[idx - 1], v[up] + 1)
                        break
        return v

    def ref_smooth(v, radius):
        v = v.copy()
        for idx in range(1, len(v)):
            if v[idx] != v[idx - 1] + 1 and v[idx - 1] != -1:
                for up in range(idx + 1, len(v)):
                    if up - idx > radius:
                        break
                    if v[up] - v[idx - 1] == up - (idx - 1):
                        v[idx - 1:up + 1] = np.arange(v[idx - 1], v[up] + 1)
                        break
        return v

    rng = np.random.default_rng(11)
    for trial in range(200):
        n = int(rng.integers(5, 400))
        v = np.arange(n, dtype=np.int32) + int(rng.integers(0, 50))
        for _ in range(int(rng.integers(0, 12))):      # holes, jumps and noise
            i = int(rng.integers(0, n))
            j = min(n, i + int(rng.integers(1, 30)))
            kind = rng.integers(0, 3)
            if kind == 0:
                v[i:j] = -1
            elif kind == 1:
                v[i:] += int(rng.integers(1, 40))
            else:
                v[i:j] = rng.integers(0, 500, j - i)
        for width in (3, 20, 300):
            a = v.copy()
            B.fillGaps(a, width)
            assert np.array_equal(a, ref_fill(v, width)), (trial, width)
            b = v.copy()
            B.smoothEdges(b, width)
            assert np.array_equal(b, ref_smooth(v, width)), (trial, width)


@needs_bin
def test_python_layer_reproduces_the_reference_lut(tmp_path):
    """reference binary -> XMFA -> our parser + walk + healing == the LUT the reference's own Python produced"""
    from mauve_py_b200 import buildindex as B
    lut, meta = _golden()
    fas = _fastas(tmp_path)
    subprocess.check_call([BINARY, "--output=mds42.xmfa", os.path.basename(fas[0]), os.path.basename(fas[1])], cwd=str(tmp_path),
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    xmfa = os.path.join(str(tmp_path), "mds42.xmfa")
    body = b"".join(b" ".join(l.split()[:3]) + b"\n" if l.startswith(b">") else l for l in open(xmfa, "rb") if not l.startswith(b"#"))
    assert hashlib.sha1(body).hexdigest() == meta["xmfa_body_sha1"]
    got = B.lut_from_xmfa(xmfa, B.getSeqFromFile(fas[0]), B.getSeqFromFile(fas[1]))
    assert got.dtype == np.int32 and np.array_equal(got, lut)


@needs_bin
def test_buildindex_flow_with_the_oracle_list(tmp_path, monkeypatch, orc):
    """everything of buildIndex() except the device call, on the CPU: with the match list the oracle computes (= the list the GPU
    returns, tests/test_gpu_parity.py::test_mums_mds42) handed to the unmodified binary through --match-input, the LUT is the
    reference's.  The product itself never falls back like this: the substitution is made here, by the test."""
    from mauve_py_b200 import buildindex as B
    lut, meta = _golden()
    fas = _fastas(tmp_path)
    bindir = os.path.join(str(tmp_path), "bin")
    os.makedirs(bindir)
    os.symlink(BINARY, os.path.join(bindir, "progressiveMauveStatic"))
    monkeypatch.setenv("MAUVE_DIR", bindir)
    monkeypatch.setattr(B.libmems, "find_mums", lambda a, b, seed, rule=0: orc.find_mums(a, b, seed, rule))
    import _oracle
    olib = _oracle.oracle()

    class OracleSML:   # stands in for the device-built DNAMemorySML: same positions (ties position-ascending), same packed sequence
        def Create(self, seq, seed):
            self.n, self.seed = len(seq), seed
            self.pos, _ = orc.sml_build(seq, seed)
            self.packed = np.zeros(int(olib.orc_packed_words(len(seq))), dtype=np.uint32)
            olib.orc_pack(seq, len(seq), self.packed.ctypes.data)

        def WriteFile(self, path):
            B.libmems.write_sslist(path, self.n, self.seed, self.packed, self.pos)

    monkeypatch.setattr(B.libmems, "DNAMemorySML", OracleSML)
    got = B.buildIndex(fas[0], fas[1])
    assert np.array_equal(got, lut)
    assert not os.path.exists(fas[0] + ".sslist") and not os.path.exists(fas[1] + ".sslist")
    # a mer beyond MER_REPEAT_LIMIT (stats[3] != 0): the anchors are NOT handed over (the reference's own list may then carry rows
    # that depend on std::sort's tie order); the binary finds them itself from the lists written for it, and the LUT is still its own
    seen = []
    real_run = B.runMauve

    def spy(files, flags):
        seen.append(dict(flags))
        return real_run(files, flags)

    def flagged(a, b, seed, rule=0):
        rows, stats = orc.find_mums(a, b, seed, rule)
        stats = np.array(stats, dtype=np.uint64)
        stats[3] = 1
        return rows, stats

    monkeypatch.setattr(B, "runMauve", spy)
    monkeypatch.setattr(B.libmems, "find_mums", flagged)
    got = B.buildIndex(fas[0], fas[1])
    assert len(seen) == 1 and "--match-input" not in seen[0] and np.array_equal(got, lut)
    monkeypatch.setattr(B, "runMauve", real_run)
    monkeypatch.delenv("MAUVE_DIR")
    with pytest.raises(IOError):
        B.buildIndex(fas[0], fas[1])


@needs_bin
@pytest.mark.gpu
def test_buildindex_dropin_mds42(tmp_path, monkeypatch):
    """BASELINE config 1 / north_star target: bit-exact MDS42 buildIndex LUT with the anchoring stage on the GPU"""
    import mauve_py_b200 as mp
    from mauve_py_b200._capi import check
    check(mp.lib().mcu_init(0))
    lut, meta = _golden()
    fas = _fastas(tmp_path)
    bindir = os.path.join(str(tmp_path), "bin")
    os.makedirs(bindir)
    os.symlink(BINARY, os.path.join(bindir, "progressiveMauveStatic"))
    monkeypatch.setenv("MAUVE_DIR", bindir)
    got = mp.buildIndex(fas[0], fas[1])
    assert got.dtype == np.int32 and got.shape == lut.shape
    assert np.array_equal(got, lut)
    assert not os.path.exists(fas[0] + ".sslist")   # cleaned up like the reference does


@needs_bin
def test_sslist_files_are_loaded_by_the_unmodified_binary(tmp_path, orc):
    """a `<fasta>.sslist` written by mauve_py_b200.libmems.write_sslist (here from the oracle's sorted mer list,
    whose tie order -- position-ascending, like the device's -- differs from the reference's std::sort) is accepted by
    MatchList::LoadSMLs instead of being rebuilt, and the alignment that follows is byte-identical: no consumer depends on the
    order inside equal-mer runs."""
    import ctypes as C
    import _oracle
    import mauve_py_b200 as mp
    from mauve_py_b200 import synth
    a, b = synth.small_pair(150000, seed=21, snp=0.02, n_inv=2)
    d = str(tmp_path)
    for name, s in (("a", a), ("b", b)):
        with open(os.path.join(d, name + ".fa"), "wb") as f:
            f.write(b">" + name.encode() + b"\n" + b"\n".join(s[i:i + 80] for i in range(0, len(s), 80)) + b"\n")

    def run(tag):
        r = subprocess.run([BINARY, "--output=%s.xmfa" % tag, "a.fa", "b.fa"], cwd=d, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-300:]
        body = b"".join(l for l in open(os.path.join(d, tag + ".xmfa"), "rb") if not l.startswith(b"#"))
        return r.stdout, body

    log0, plain = run("plain")
    assert "Creating sorted mer list" in log0
    ref_files = {n: open(os.path.join(d, n + ".fa.sslist"), "rb").read() for n in ("a", "b")}
    seed = mp.getSeed(mp.getDefaultSeedWeight((len(a) + len(b)) // 2), mp.CODING_SEED)
    lib = _oracle.oracle()
    for name, s in (("a", a), ("b", b)):
        os.remove(os.path.join(d, name + ".fa.sslist"))
        pos, mer = orc.sml_build(s, seed)
        packed = np.zeros(int(lib.orc_packed_words(len(s))), dtype=np.uint32)
        lib.orc_pack(s, len(s), packed.ctypes.data)
        mp.libmems.write_sslist(os.path.join(d, name + ".fa.sslist"), len(s), seed, packed, pos)
        ours = open(os.path.join(d, name + ".fa.sslist"), "rb").read()
        theirs = ref_files[name]
        H = mp.libmems.SML_HEADER_BYTES
        assert len(ours) == len(theirs) and ours[:36] == theirs[:36]            # version .. unique_mers
        assert ours[44:300] == theirs[44:300]                                   # circular flag + translation table
        nseq = packed.size * 4
        core = ((2 * len(s) + 31) // 32) * 4                                    # the reference leaves its two pad words uninitialised
        assert ours[H:H + core] == theirs[H:H + core]                           # the 2-bit sequence
        mine, ref_pos = np.frombuffer(ours[H + nseq:], dtype=np.uint32), np.frombuffer(theirs[H + nseq:], dtype=np.uint32)
        assert np.array_equal(np.sort(mine), np.sort(ref_pos)) and not np.array_equal(mine, ref_pos)   # same list up to tie order
    log1, loaded = run("loaded")
    assert "Sorted mer list loaded successfully" in log1 and "Creating sorted mer list" not in log1
    assert loaded == plain


@needs_bin
def test_sslist_reader_on_the_reference_binarys_own_files(tmp_path, orc):
    """SURVEY.md 8f-3, the reading side: mauve_py_b200.libmems.read_sslist (= FileSML::LoadFile2, LM/FileSML.cpp:120-195) on the
    `.sslist` the UNMODIFIED binary leaves next to its FASTA input: header fields, the 2-bit sequence and the sorted positions are
    the oracle's sorted mer list (same mer at every rank, same position multiset inside every equal-mer run: SURVEY.md 8a-4);
    LoadFile2's return codes for files cut short; write_sslist -> read_sslist round trip."""
    import _oracle
    import mauve_py_b200 as mp
    from mauve_py_b200 import synth
    L_ = mp.libmems
    a, b = synth.small_pair(60000, seed=33, snp=0.03, n_inv=1)
    d = str(tmp_path)
    for name, s in (("a", a), ("b", b)):
        with open(os.path.join(d, name + ".fa"), "wb") as f:
            f.write(b">" + name.encode() + b"\n" + b"\n".join(s[i:i + 80] for i in range(0, len(s), 80)) + b"\n")
    r = subprocess.run([BINARY, "--output=x.xmfa", "a.fa", "b.fa"], cwd=d, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-300:]
    seed = mp.getSeed(mp.getDefaultSeedWeight((len(a) + len(b)) // 2), mp.CODING_SEED)
    lib = _oracle.oracle()
    for name, s in (("a", a), ("b", b)):
        path = os.path.join(d, name + ".fa.sslist")
        code, h, packed, pos = L_.read_sslist(path)
        assert code == 0
        assert (h["version"], h["alphabet_bits"], h["seed"], h["length"], h["circular"]) == (L_.SML_FORMAT_VERSION, 2, seed, len(s), 0)
        assert h["seed_length"] == mp.getSeedLength(seed) and h["seed_weight"] == mp.getSeedWeight(seed)
        want = np.zeros(int(lib.orc_packed_words(len(s))), dtype=np.uint32)
        lib.orc_pack(s, len(s), want.ctypes.data)
        core = (2 * len(s) + 31) // 32   # the reference leaves its two pad words uninitialised
        assert packed.size == want.size and np.array_equal(packed[:core], want[:core])
        opos, omer = orc.sml_build(s, seed)
        assert pos.size == opos.size == len(s) - h["seed_length"] + 1
        mer_at = np.empty(len(s), dtype=np.uint64)
        mer_at[opos] = omer                                  # canonical seed mer of every position, from the oracle
        assert np.array_equal(mer_at[pos], omer)             # the reference's file: same mer at every rank
        key = np.lexsort((pos, mer_at[pos]))                 # ties position-ascending == the oracle's (and the device's) order
        assert np.array_equal(pos[key], opos)
        sml = mp.DNAMemorySML()
        assert sml.LoadFile(path) == 0 and sml.SMLLength() == pos.size and sml.Seed() == seed and sml.Length() == len(s)
        # files cut short: LoadFile2's codes
        raw = open(path, "rb").read()
        H = L_.SML_HEADER_BYTES
        cases = {2: raw[:H - 1], 4: raw[:H + 4 * packed.size - 1], 5: raw[:-1], 3: b"\x04" + raw[1:]}
        for want_code, data in cases.items():
            q = os.path.join(d, "cut%d.sslist" % want_code)
            open(q, "wb").write(data)
            assert L_.read_sslist(q)[0] == want_code
            assert mp.DNAMemorySML().LoadFile(q) == want_code
        assert L_.read_sslist(os.path.join(d, "absent.sslist"))[0] == 1
        # round trip of our writer
        q = os.path.join(d, "rt.sslist")
        L_.write_sslist(q, len(s), seed, want, opos)
        code, h2, packed2, pos2 = L_.read_sslist(q)
        assert code == 0 and np.array_equal(packed2, want) and np.array_equal(pos2, opos) and h2["seed"] == seed and h2["little_endian"] == 1
        assert np.array_equal(h2["translation_table"], h["translation_table"])


# ---- progressiveMauve_cuda / progressiveMauve_cuda_mh: the reference binary with link-time seams (adapters/seams) ---------------
CUDA_BINARY = os.path.join(REF_DIR, "progressiveMauve_cuda")         # gapped DP of every window on the device
CUDA_MH_BINARY = os.path.join(REF_DIR, "progressiveMauve_cuda_mh")   # + MemHash::FindMatches of two genomes on the device
CUDA_ALL_BINARY = os.path.join(REF_DIR, "progressiveMauve_cuda_all")  # + the DP of RefineW's windows prefetched on the device
needs_cuda_bin = pytest.mark.skipif(not all(os.path.exists(b) for b in (CUDA_BINARY, CUDA_MH_BINARY, CUDA_ALL_BINARY)),
                                    reason="oracle/_ref/progressiveMauve_cuda[_mh|_all] not built (needs /root/reference at build time)")


def _xmfa_body_sha1(xmfa):
    body = b"".join(b" ".join(l.split()[:3]) + b"\n" if l.startswith(b">") else l for l in open(xmfa, "rb") if not l.startswith(b"#"))
    return hashlib.sha1(body).hexdigest()


def _align(binary, d, a, b, out, env=None):
    for f in os.listdir(d):
        if f.endswith(".sslist"):
            os.remove(os.path.join(d, f))
    return subprocess.run([binary, "--output=" + out, a, b], cwd=d, capture_output=True, text=True, env=env)


def _spiked_pair(d):
    """a 80 kbp pair with an inversion, single N and runs of N in both genomes: a.fa, b.fa in directory d.  The N columns make some DP
    ranges / refine windows fall outside the integer kernel's form, so the seams' fall-back to the reference's own code is exercised"""
    from mauve_py_b200 import synth
    a, b = synth.small_pair(80000, seed=77, snp=0.03, n_inv=1)
    rng = np.random.default_rng(5)
    for name, s_ in (("a", a), ("b", b)):
        s_ = bytearray(s_)
        for i in rng.integers(0, len(s_), 60):
            s_[i] = ord("N")
        for i in rng.integers(0, len(s_) - 30, 5):
            s_[i:i + 30] = b"N" * 30
        with open(os.path.join(d, name + ".fa"), "wb") as f:
            f.write(b">" + name.encode() + b"\n" + b"\n".join(bytes(s_[i:i + 70]) for i in range(0, len(s_), 70)) + b"\n")


def _seam_counts(stderr):
    out = {}
    for l in stderr.splitlines():
        if l.split(" seam:")[0] in ("AnchoredProfileProfile", "FindAnchorColsPP", "MemHash::FindMatches", "RefineW", "SeedOccurrenceList::construct", "FileSML::Create"):
            out[l.split(" seam:")[0]] = [int(x) for x in l.replace(",", "").split() if x.isdigit()]
        if l.startswith("EliminateOverlaps_v2 seam:") or l.startswith("IdentifyBreakpoints seam:"):
            out[l.split(" seam:")[0]] = [int(x) for x in l.replace(",", "").replace("(", " ").split() if x.isdigit()]
        if l.startswith("run() seam (HomologyHMM):"):   # strings on the device, their columns, strings left to the reference's run()
            out["run"] = [int(x) for x in l.replace(",", "").replace("(", " ").split() if x.isdigit()]
    return out


@needs_cuda_bin
def test_seams_host_code_inside_the_reference_binary(tmp_path):
    """The link-time seams (mauve_py_b200/adapters/seams: all DP ranges of a window -> one CudaGlobalAlignBatch call;
    MemHash::FindMatches of two genomes -> mcu_find_mums; the DP of all windows of a RefineW call prefetched in one mcu_nw_batch
    call; FileSML::Create -> mcu_sml_build; SeedOccurrenceList::construct -> mcu_sol_build) inside the unmodified reference objects align the MDS42 pair to the byte-identical XMFA, with EVERY ONE of the run's
    61,773 gapped-DP calls answered by the device entry point.  Here, without a GPU, the device entry points are answered by the CPU restatement through an LD_PRELOAD
    stub (tests/_stub): this checks the seams' own host code -- range collection, profile order, path -> PWPath, output assembly,
    sequence extraction from progressiveMauve's gnRAWSequence objects, Match construction; the GPU suite runs the same binaries
    against the real library.  Without the stub the binaries stop with the library's error: no CPU fallback behind the seams."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present: the GPU suite runs these binaries against the real library")
    import _emu
    from mauve_py_b200 import synth
    d = str(tmp_path)
    _spiked_pair(d)
    r = _align(CUDA_BINARY, d, "a.fa", "b.fa", "x.xmfa")
    assert r.returncode == 3 and "no usable CUDA device" in (r.stdout + r.stderr)
    r = _align(CUDA_MH_BINARY, d, "a.fa", "b.fa", "x.xmfa")
    assert r.returncode != 0 and "no usable CUDA device" in (r.stdout + r.stderr)
    env = dict(os.environ, LD_PRELOAD=_emu.stub_library(), MAUVE_CUDA_SEAM_REPORT="1", MAUVE_CUDA_GAP_SEAM="1", MAUVE_CUDA_SOL_SEAM="1")
    assert _align(BINARY, d, "a.fa", "b.fa", "ref.xmfa").returncode == 0
    # default: ranges and windows with N columns go to the float wavefront kernel's entry point (mcu_nw_batch_wild): nothing is left
    # to the reference's NWSmall, and the alignment is still the reference's
    for binary in (CUDA_BINARY, CUDA_ALL_BINARY):
        r = _align(binary, d, "a.fa", "b.fa", "seam.xmfa", env)
        assert r.returncode == 0 and _xmfa_body_sha1(os.path.join(d, "seam.xmfa")) == _xmfa_body_sha1(os.path.join(d, "ref.xmfa")), r.stderr[-500:]
        c = _seam_counts(r.stderr)
        calls, ranges, device = c["AnchoredProfileProfile"]
        assert calls >= 1 and 10 < device == ranges
        assert "mcu_nw_batch_wild problems" in r.stderr
        if binary == CUDA_ALL_BINARY:
            assert c["RefineW"][2] > 10 and c["RefineW"][3] == 0
            assert c["MemHash::FindMatches"][0] > 10 and c["FileSML::Create"] == [2, 0] and c["SeedOccurrenceList::construct"] == [2, 0]
    # MAUVE_CUDA_WILD=0 (a switch for A/B runs): ranges with an N column take the reference's ProfileProfile, windows with an N are not prefetched
    env_w = dict(env, MAUVE_CUDA_WILD="0")
    r = _align(CUDA_ALL_BINARY, d, "a.fa", "b.fa", "wild.xmfa", env_w)
    assert r.returncode == 0 and _xmfa_body_sha1(os.path.join(d, "wild.xmfa")) == _xmfa_body_sha1(os.path.join(d, "ref.xmfa")), r.stderr[-500:]
    c = _seam_counts(r.stderr)
    assert c["AnchoredProfileProfile"][2] < c["AnchoredProfileProfile"][1] and c["RefineW"][3] > 0
    # BASELINE config 1 with every seam on: initial anchors, the gap searches of recursive anchoring, the DP of every window
    _lut, meta = _golden()
    fas = _fastas(tmp_path)
    r = _align(CUDA_ALL_BINARY, d, os.path.basename(fas[0]), os.path.basename(fas[1]), "cuda.xmfa", env)
    assert r.returncode == 0, r.stderr[-500:]
    assert _xmfa_body_sha1(os.path.join(d, "cuda.xmfa")) == meta["xmfa_body_sha1"]
    c = _seam_counts(r.stderr)
    assert c["MemHash::FindMatches"][0] > 300                 # 1 initial search + the 371 gap searches
    assert c["AnchoredProfileProfile"][0] > 100 and c["AnchoredProfileProfile"][1] == c["AnchoredProfileProfile"][2] > 40000   # 165 windows, 41,806 ranges
    calls, prefetched, hits, misses = c["RefineW"]
    assert calls > 100 and hits > 19000 and misses == 0       # 19,967 windows of RefineFast: every GlobalAlign answered from the prefetch
    assert c["AnchoredProfileProfile"][2] + hits == 61773     # = all gapped-DP calls of the run (tests/golden/dp_mds42_calls.npz meta)
    assert c["SeedOccurrenceList::construct"] == [2, 0]       # both genomes' seed occurrence lists (adapters/seams/sol_seam.cpp)
    assert c["FileSML::Create"] == [2, 0]                     # both `.sslist` files (adapters/seams/filesml_seam.cpp)
    assert c["run"][0] >= 1 and c["run"][1] > 3_900_000 and c["run"][2] == 0   # the backbone HMM: one string as long as the alignment (adapters/seams/hmm_seam.cpp)
    # the step after every match list (adapters/seams/lcb_seam.cpp): overlaps of the initial list (25 tied keys) and of the gap lists that hold two matches or more, the LCBs
    assert c["EliminateOverlaps_v2"][0] > 50 and c["EliminateOverlaps_v2"][1] >= 25 and c["EliminateOverlaps_v2"][2] == 0
    assert c["IdentifyBreakpoints"][0] >= 1 and c["IdentifyBreakpoints"][1] == 0
    # the anchor columns of every window (adapters/CudaAnchorCols.h inside the AnchoredProfileProfile seam): all 165 windows, and the
    # ranges between the columns are the 41,806 above
    assert c["FindAnchorColsPP"][0] == c["AnchoredProfileProfile"][0] and c["FindAnchorColsPP"][1] > 40000


@needs_cuda_bin
@pytest.mark.parametrize("option", ["--seed-family", "--collinear", "--skip-refinement", "--solid-seeds"])
def test_seam_binary_under_other_command_lines(tmp_path, option):
    """other paths through the aligner (a family of three seeds through UniqueMatchFinder, collinear genomes, no refinement, solid
    seeds): the seam binary's alignment equals the reference binary's.  CPU: device calls answered through the stub; with a GPU:
    the real library."""
    import torch
    d = str(tmp_path)
    _spiked_pair(d)
    env = dict(os.environ, MAUVE_CUDA_SEAM_REPORT="1", MAUVE_CUDA_SOL_SEAM="1", MAUVE_CUDA_WILD="1")
    if not torch.cuda.is_available():
        import _emu
        env["LD_PRELOAD"] = _emu.stub_library()

    def run(binary, out, e=None):
        for f in os.listdir(d):
            if f.endswith(".sslist"):
                os.remove(os.path.join(d, f))
        return subprocess.run([binary, option, "--output=" + out, "a.fa", "b.fa"], cwd=d, capture_output=True, text=True, env=e)

    assert run(BINARY, "ref.xmfa").returncode == 0
    r = run(CUDA_ALL_BINARY, "seam.xmfa", env)
    assert r.returncode == 0, r.stderr[-500:]
    assert _xmfa_body_sha1(os.path.join(d, "seam.xmfa")) == _xmfa_body_sha1(os.path.join(d, "ref.xmfa"))
    c = _seam_counts(r.stderr)
    assert c["AnchoredProfileProfile"][1] == c["AnchoredProfileProfile"][2] > 100
    if option == "--seed-family":
        assert c["MemHash::FindMatches"][0] == 0          # UniqueMatchFinder and the hash-table readers stay with the reference's code


@needs_cuda_bin
def test_seam_binary_three_genomes(tmp_path):
    """three genomes: the pairwise parts (gap searches of the pairwise recursion, DP ranges between two single sequences, one seed
    occurrence list and one `.sslist` per genome) go through the C ABI, everything that involves profiles of several sequences or
    three-way match finding stays with the reference's code; the alignment equals the reference binary's"""
    import torch
    from mauve_py_b200 import synth
    d = str(tmp_path)
    a, b = synth.small_pair(50000, seed=91, snp=0.03, n_inv=1)
    c = synth.snps(np.frombuffer(a, dtype=np.uint8), 0.04, synth.rng_for(92)).tobytes()
    for name, s_ in (("g1", a), ("g2", b), ("g3", c)):
        with open(os.path.join(d, name + ".fa"), "wb") as f:
            f.write(b">" + name.encode() + b"\n" + b"\n".join(s_[i:i + 70] for i in range(0, len(s_), 70)) + b"\n")
    env = dict(os.environ, MAUVE_CUDA_SEAM_REPORT="1", MAUVE_CUDA_GAP_SEAM="1", MAUVE_CUDA_SOL_SEAM="1")
    if not torch.cuda.is_available():
        import _emu
        env["LD_PRELOAD"] = _emu.stub_library()

    def run(binary, out, e=None):
        for f in os.listdir(d):
            if f.endswith(".sslist"):
                os.remove(os.path.join(d, f))
        return subprocess.run([binary, "--output=" + out, "g1.fa", "g2.fa", "g3.fa"], cwd=d, capture_output=True, text=True, env=e)

    assert run(BINARY, "ref.xmfa").returncode == 0
    r = run(CUDA_ALL_BINARY, "seam.xmfa", env)
    assert r.returncode == 0, r.stderr[-500:]
    assert _xmfa_body_sha1(os.path.join(d, "seam.xmfa")) == _xmfa_body_sha1(os.path.join(d, "ref.xmfa"))
    cnt = _seam_counts(r.stderr)
    assert cnt["FileSML::Create"] == [3, 0] and cnt["SeedOccurrenceList::construct"] == [3, 0]
    assert 0 < cnt["AnchoredProfileProfile"][2] < cnt["AnchoredProfileProfile"][1]     # ranges between multi-sequence profiles: reference
    assert cnt["MemHash::FindMatches"][0] > 10 and cnt["MemHash::FindMatches"][1] > 0  # the three-way searches: reference


@needs_cuda_bin
@pytest.mark.gpu
@pytest.mark.parametrize("binary,gap_seam,sol_seam", [(CUDA_BINARY, "0", "0"), (CUDA_MH_BINARY, "1", "0"),
                                                      (CUDA_ALL_BINARY, "1", "0"), (CUDA_ALL_BINARY, "1", "1")],
                         ids=["dp", "dp+pmf+gaps", "dp+pmf+gaps+refine+sml", "dp+pmf+gaps+refine+sml+sol"])
def test_buildindex_with_the_seam_binaries_mds42(tmp_path, monkeypatch, binary, gap_seam, sol_seam):
    """mauve_py_b200.buildIndex with a seam binary as $MAUVE_DIR/progressiveMauveStatic: sorted mer lists, anchors AND the gapped DP
    of every window (and, last case, the gap searches of recursive anchoring) on the device; the LUT is the reference's.  Then the
    binary on its own, from the FASTA files: byte-identical XMFA."""
    import time
    import mauve_py_b200 as mp
    from mauve_py_b200._capi import check
    check(mp.lib().mcu_init(0))
    lut, meta = _golden()
    fas = _fastas(tmp_path)
    bindir = os.path.join(str(tmp_path), "bin")
    os.makedirs(bindir)
    os.symlink(binary, os.path.join(bindir, "progressiveMauveStatic"))
    monkeypatch.setenv("MAUVE_DIR", bindir)
    monkeypatch.setenv("MAUVE_CUDA_GAP_SEAM", gap_seam)
    monkeypatch.setenv("MAUVE_CUDA_SOL_SEAM", sol_seam)
    t0 = time.time()
    got = mp.buildIndex(fas[0], fas[1])
    t1 = time.time()
    assert np.array_equal(got, lut)
    env = dict(os.environ, MAUVE_CUDA_SEAM_REPORT="1")
    r = _align(binary, str(tmp_path), os.path.basename(fas[0]), os.path.basename(fas[1]), "cuda.xmfa", env)
    t2 = time.time()
    assert r.returncode == 0, r.stderr[-500:]
    assert _xmfa_body_sha1(os.path.join(str(tmp_path), "cuda.xmfa")) == meta["xmfa_body_sha1"]
    c = _seam_counts(r.stderr)
    assert c["AnchoredProfileProfile"][1] == c["AnchoredProfileProfile"][2] > 40000
    assert c["FindAnchorColsPP"][0] == c["AnchoredProfileProfile"][0] and c["FindAnchorColsPP"][1] > 40000   # every window's anchor columns on the device
    if binary != CUDA_BINARY:
        assert c["MemHash::FindMatches"][0] >= (300 if gap_seam == "1" else 1)
    if binary == CUDA_ALL_BINARY:
        assert c["RefineW"][3] == 0 and c["AnchoredProfileProfile"][2] + c["RefineW"][2] == 61773
        assert c["SeedOccurrenceList::construct"] == ([2, 0] if sol_seam == "1" else [0, 2])
        assert c["FileSML::Create"] == [2, 0]
        assert c["run"][0] >= 1 and c["run"][1] > 3_900_000 and c["run"][2] == 0
        assert c["EliminateOverlaps_v2"][0] > (50 if gap_seam == "1" else 0) and c["EliminateOverlaps_v2"][2] == 0 and c["IdentifyBreakpoints"] == [1, 0]
    # a pair with N columns: part of the DP falls back to the reference's code, the alignment stays the reference's
    d2 = os.path.join(str(tmp_path), "spiked")
    os.makedirs(d2)
    _spiked_pair(d2)
    assert _align(BINARY, d2, "a.fa", "b.fa", "ref.xmfa").returncode == 0
    r = _align(binary, d2, "a.fa", "b.fa", "seam.xmfa", env)
    assert r.returncode == 0 and _xmfa_body_sha1(os.path.join(d2, "seam.xmfa")) == _xmfa_body_sha1(os.path.join(d2, "ref.xmfa")), r.stderr[-500:]
    print("buildIndex %.1f s, standalone binary %.1f s (%s, gap seam %s, sol seam %s)" % (t1 - t0, t2 - t1, os.path.basename(binary), gap_seam, sol_seam))


@needs_cuda_bin
@pytest.mark.gpu
def test_seam_binary_config2_pair(tmp_path):
    """BASELINE config 2 (synthetic 5 Mbp bacterial pair: SNPs, codon recoding, indels) through the binary with every seam on: the XMFA
    equals the reference binary's.  (With the device calls answered by the CPU restatement this was checked in the build container:
    49,808 DP ranges, 25,149 refine windows, 2,212 match-finder calls, sha1 of the XMFA body 3fd365f4...)"""
    from mauve_py_b200 import synth
    d = str(tmp_path)
    a, b = synth.config2_pair()
    for name, s_ in (("a", a.tobytes()), ("b", b.tobytes())):
        with open(os.path.join(d, name + ".fa"), "wb") as f:
            f.write(b">" + name.encode() + b"\n" + b"\n".join(s_[i:i + 80] for i in range(0, len(s_), 80)) + b"\n")
    assert _align(BINARY, d, "a.fa", "b.fa", "ref.xmfa").returncode == 0
    env = dict(os.environ, MAUVE_CUDA_SEAM_REPORT="1", MAUVE_CUDA_GAP_SEAM="1", MAUVE_CUDA_SOL_SEAM="0")
    r = _align(CUDA_ALL_BINARY, d, "a.fa", "b.fa", "seam.xmfa", env)
    assert r.returncode == 0, r.stderr[-500:]
    want = _xmfa_body_sha1(os.path.join(d, "ref.xmfa"))
    assert want.startswith("3fd365f4") and _xmfa_body_sha1(os.path.join(d, "seam.xmfa")) == want
    c = _seam_counts(r.stderr)
    assert c["AnchoredProfileProfile"][1] == c["AnchoredProfileProfile"][2] > 40000 and c["RefineW"][3] == 0 and c["MemHash::FindMatches"][0] > 2000


@needs_cuda_bin
@pytest.mark.gpu
def test_seam_binary_wildcards_on_device(tmp_path):
    """MAUVE_CUDA_WILD=1: the DP ranges and refine windows that contain N columns go to mcu_nw_batch_wild (the reference's float
    arithmetic on the device) instead of the reference's NWSmall; every DP of the run is then on the device and the XMFA is the
    reference binary's"""
    d = str(tmp_path)
    _spiked_pair(d)
    assert _align(BINARY, d, "a.fa", "b.fa", "ref.xmfa").returncode == 0
    env = dict(os.environ, MAUVE_CUDA_SEAM_REPORT="1", MAUVE_CUDA_GAP_SEAM="1", MAUVE_CUDA_SOL_SEAM="0", MAUVE_CUDA_WILD="1")
    r = _align(CUDA_ALL_BINARY, d, "a.fa", "b.fa", "wild.xmfa", env)
    assert r.returncode == 0, r.stderr[-500:]
    assert _xmfa_body_sha1(os.path.join(d, "wild.xmfa")) == _xmfa_body_sha1(os.path.join(d, "ref.xmfa"))
    c = _seam_counts(r.stderr)
    assert c["AnchoredProfileProfile"][1] == c["AnchoredProfileProfile"][2] > 100 and c["RefineW"][2] > 100 and c["RefineW"][3] == 0
