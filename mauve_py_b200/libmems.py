"""Host-side mirror of the libMems / libMUSCLE / HomologyHMM interfaces of the anchoring path.

Same names, argument meaning and error behaviour as the reference classes the path's callers use
(SURVEY.md 8b), so that parity tests read like the reference's own call sites:

    sml = DNAMemorySML(); sml.Create(seq, getSeed(15, CODING_SEED))          LM/MemorySML.cpp:45
    ml = MatchList(); ml.seq_table = [a, b]; ml.CreateMemorySMLs(0, CODING_SEED)   LM/MatchList.h:431
    PairwiseMatchFinder().FindMatches(ml)                                     LM/MemHash.cpp:109
    paths = GlobalAlignBatch(pairs)                                           MU/glbalign.cpp:69
    pred = run(symbols, params)                                               LM/HomologyHMM/homologymain.cc:24

Every method that computes goes through the C ABI of libmauve_cuda.so (include/mauve_cuda.h);
this module holds containers and argument marshalling only.
"""
import ctypes as C
from dataclasses import dataclass, field
from typing import List

import numpy as np

from . import _capi
from ._capi import CODING_SEED, SOLID_SEED, McuError, check, lib

__all__ = [
    "CODING_SEED", "SOLID_SEED", "McuError", "getSeed", "getSolidSeed", "getDefaultSeedWeight", "getSeedLength", "getSeedWeight",
    "bmer", "DNAMemorySML", "read_sslist", "write_sslist", "Match", "MatchList", "MemHash", "PairwiseMatchFinder", "AnchorSession", "merge_matches",
    "PWPath", "GlobalAlign", "GlobalAlignBatch", "GlobalAlignBatchWild", "Params", "hmm_params", "getAdaptedHoxdMatrixParameters", "adaptToPercentIdentity",
    "run", "run_batch", "sort_pairs", "SeedOccurrenceList", "GetPairwiseAnchorScore", "anchor_scores", "hoxd_matrix",
    "EliminateOverlaps_v2", "IdentifyBreakpoints", "ComputeLCBs_v2", "sml_build_shard", "sml_build_sharded",
    "AnchorParams", "FindAnchorColsPP", "FindAnchorColsPP_batch",
]


def _buf(x):
    """bytes / bytearray / numpy uint8 -> (address, length, keepalive)"""
    if isinstance(x, (bytes, bytearray)):
        arr = np.frombuffer(x, dtype=np.uint8)
    else:
        arr = np.ascontiguousarray(x, dtype=np.uint8)
    return arr.ctypes.data, arr.size, arr


# ---- LM/SeedMasks.h ---------------------------------------------------------------------------
def getSeed(weight: int, seed_rank: int = 0) -> int:
    return int(lib().mcu_get_seed(weight, seed_rank))


def getSolidSeed(weight: int) -> int:
    return int(lib().mcu_get_seed(weight, SOLID_SEED))


def getDefaultSeedWeight(avg_sequence_length: int) -> int:
    return int(lib().mcu_default_seed_weight(int(avg_sequence_length)))


def getSeedLength(seed: int) -> int:
    return int(lib().mcu_seed_length(seed))


def getSeedWeight(seed: int) -> int:
    return int(lib().mcu_seed_weight(seed))


# ---- LM/SortedMerList.h, LM/MemorySML.h, LM/DNAMemorySML.h ---------------------------------------
@dataclass
class bmer:
    position: int
    mer: int


class DNAMemorySML:
    """mems::DNAMemorySML: the sorted mer list of one genome, built on the GPU."""

    def __init__(self):
        self.Clear()

    def Clear(self):
        self._pos = np.zeros(0, dtype=np.uint32)
        self._mer = np.zeros(0, dtype=np.uint64)
        self._packed = np.zeros(0, dtype=np.uint32)
        self._seed = 0
        self._length = 0
        self._seq = np.zeros(0, dtype=np.uint8)  # the sequence the list was built from (SortedMerList keeps it 2-bit packed)

    def Create(self, seq, seed: int):
        """SortedMerList::Create + FillDnaSeedSML + sort (LM/MemorySML.cpp:45-60)."""
        addr, n, keep = _buf(seq)
        L = getSeedLength(seed)
        m = max(n - L + 1, 0) if L else 0
        pos = np.empty(max(m, 1), dtype=np.uint32)
        mer = np.empty(max(m, 1), dtype=np.uint64)
        packed = np.empty((2 * n) // 32 + (1 if (2 * n) % 32 else 0) + 2, dtype=np.uint32)
        out_len = C.c_uint64(0)
        check(lib().mcu_sml_build(addr, n, seed, pos.ctypes.data, mer.ctypes.data, packed.ctypes.data, C.byref(out_len)))
        k = out_len.value
        self._pos, self._mer, self._packed = pos[:k], mer[:k], packed
        self._seed, self._length, self._seq = seed, n, keep

    def Length(self):
        return self._length

    def SMLLength(self):
        return int(self._pos.size)

    def Seed(self):
        return self._seed

    def SeedLength(self):
        return getSeedLength(self._seed)

    def SeedWeight(self):
        return getSeedWeight(self._seed)

    def GetSeedMask(self):
        w = self.SeedWeight()
        return ((1 << 64) - 1) ^ ((1 << (64 - 2 * w)) - 1) if w else 0

    def Read(self, size: int, offset: int):
        """MemorySML::Read (LM/MemorySML.cpp:62-82): (positions, mers) of ranks [offset, offset+size)."""
        end = min(offset + size, self.SMLLength())
        return self._pos[offset:end], self._mer[offset:end]

    def __getitem__(self, index: int) -> bmer:
        return bmer(int(self._pos[index]), int(self._mer[index]))

    # ---- host-side accessors the reference's callers use besides Read; plain integer code on the packed sequence ----
    def _forward_mer(self, position: int) -> int:
        """SortedMerList::GetMer (LM/SortedMerList.cpp:321-342): the SeedLength() bases at `position`, left-aligned in 64 bits"""
        w, bit = (2 * position) // 32, (2 * position) % 32
        seq = self._packed
        x = (int(seq[w]) << 32) | int(seq[w + 1])
        if bit:
            x = ((x << bit) | (int(seq[w + 2]) >> (32 - bit))) & 0xFFFFFFFFFFFFFFFF
        L = self.SeedLength()
        return x & (((1 << 64) - 1) ^ ((1 << (64 - 2 * L)) - 1))

    @staticmethod
    def _revcomp_mer(mer: int, length: int) -> int:
        """SortedMerList::RevCompMer (:597-614): complement, reverse the 2-bit groups, left-align, strand flag in bit 0"""
        x = (~mer & 0xFFFFFFFFFFFFFFFF) >> (64 - 2 * length)
        rc = 0
        for _ in range(length):
            rc = (rc << 2) | (x & 3)
            x >>= 2
        return ((rc << (64 - 2 * length)) & 0xFFFFFFFFFFFFFFFF) | 1

    def GetMer(self, position: int) -> int:
        """DNAMemorySML::GetMer = SortedMerList::GetDnaMer (LM/DNAMemorySML.cpp:35-37, LM/SortedMerList.cpp:581-593): the smaller of the
        contiguous SeedLength()-mer at `position` and its reverse complement"""
        fwd = self._forward_mer(position)
        rc = self._revcomp_mer(fwd, self.SeedLength())
        return fwd if fwd < rc else rc

    def _forward_seed_mer(self, position: int) -> int:
        """SortedMerList::GetSeedMer (:726-762): the bases under the 1s of the seed pattern (pattern MSB = first base), left-aligned"""
        mer, L, w = self._forward_mer(position), self.SeedLength(), self.SeedWeight()
        out = 0
        for i in range(L):
            if (self._seed >> (L - 1 - i)) & 1:
                out = (out << 2) | ((mer >> (62 - 2 * i)) & 3)
        return (out << (64 - 2 * w)) & 0xFFFFFFFFFFFFFFFF

    def GetDnaSeedMer(self, position: int) -> int:
        """SortedMerList::GetDnaSeedMer (:764-769): the smaller of the forward seed mer and its reverse complement (strand flag in
        bit 0); a tie keeps the forward one"""
        fwd = self._forward_seed_mer(position)
        rc = self._revcomp_mer(fwd, self.SeedWeight())
        return fwd if fwd < rc else rc

    def GetSeedMer(self, position: int) -> int:
        """DNAMemorySML::GetSeedMer = GetDnaSeedMer (LM/DNAMemorySML.cpp:39-41): what ExtendMatch and operator[] compare"""
        return self.GetDnaSeedMer(position)

    def FindMer(self, query_mer: int):
        """SortedMerList::FindMer (:170-179) over bsearch (:380-394) -> (found, rank): the rank the recursion stops at (where the mer
        would be when it is absent), found = the mer AT that rank equals the query (strand flag included, as operator[] returns it)"""
        last = self._length
        L = self.SeedLength()
        if last == 0 or last < L:
            return False, 0
        start, end = 0, last - L
        while True:
            middle = (start + end) // 2
            m = int(self._mer[middle])
            if m == query_mer:
                break
            if m < query_mer and middle < end:
                start = middle + 1
            elif m > query_mer and start < middle:
                end = middle - 1
            else:
                break
        return int(self._mer[middle]) == query_mer, middle

    def Clone(self):
        """gnClone::Clone: an independent copy"""
        c = DNAMemorySML()
        c._pos, c._mer, c._packed = self._pos.copy(), self._mer.copy(), self._packed.copy()
        c._seed, c._length, c._seq = self._seed, self._length, self._seq
        return c

    def GetHeader(self):
        """the SMLHeader fields Create sets (LM/SortedMerList.cpp:786-824)"""
        return {"version": SML_FORMAT_VERSION, "alphabet_bits": 2, "seed": self._seed, "seed_length": self.SeedLength(),
                "seed_weight": self.SeedWeight(), "length": self._length, "unique_mers": 0xFFFFFFFF, "circular": 0}

    def positions(self):
        return self._pos

    def WriteFile(self, path):
        """leave this list on disk as `path` (normally <fasta>.sslist) in DNAFileSML's format"""
        write_sslist(path, self._length, self._seed, self._packed, self._pos)

    def LoadFile(self, path, seq=None):
        """take the list from a `.sslist` file (FileSML::LoadFile2's return code; 0 = loaded).  The file holds no mers (the reference
        recomputes them with GetSeedMer on every Read): mers() stays empty unless the list is rebuilt with Create."""
        code, header, packed, positions = read_sslist(path)
        if code != 0:
            return code
        self._pos, self._packed, self._mer = positions, packed, np.zeros(0, dtype=np.uint64)
        self._seed, self._length = int(header["seed"]), int(header["length"])
        self._seq = np.zeros(0, dtype=np.uint8) if seq is None else seq
        return 0

    def mers(self):
        return self._mer

    def packed_sequence(self):
        """SortedMerList::sequence: 2-bit packed, MSB first, two zero pad words."""
        return self._packed



# ---- LM/Match.h, LM/MatchList.h -----------------------------------------------------------------
@dataclass
class Match:
    """Ungapped match of two genomes: 1-based starts, negative start = reverse strand."""
    length: int
    starts: List[int]

    def Length(self):
        return self.length

    def Start(self, seq: int):
        return self.starts[seq]

    def Orientation(self, seq: int):
        return 0 if self.starts[seq] == 0 else (1 if self.starts[seq] > 0 else -1)


@dataclass
class MatchList:
    seq_table: list = field(default_factory=list)
    sml_table: list = field(default_factory=list)
    matches: List[Match] = field(default_factory=list)

    def CreateMemorySMLs(self, mer_size: int = 0, seed_rank: int = 0):
        """LM/MatchList.h:431-461: default weight from the average length, one DNAMemorySML per genome."""
        if mer_size == 0:
            avg = sum(len(s) for s in self.seq_table) // max(len(self.seq_table), 1)
            mer_size = getDefaultSeedWeight(avg)
        seed = getSeed(mer_size, seed_rank)
        self.sml_table = []
        for s in self.seq_table:
            sml = DNAMemorySML()
            sml.Create(s, seed)
            self.sml_table.append(sml)

    def __len__(self):
        return len(self.matches)

    def __getitem__(self, i):
        return self.matches[i]

    def as_array(self):
        """rows (length, start0, start1) in list order -- the three leading columns of WriteList (LM/MatchList.h:617-662)"""
        return np.array([[m.length, m.starts[0], m.starts[1]] for m in self.matches], dtype=np.int64).reshape(-1, 3)


SML_FORMAT_VERSION = 5            # DNAFileSML::FormatVersion, LM/DNAFileSML.h:59-62
SML_HEADER_BYTES = 2352           # sizeof(struct SMLHeader), LM/SortedMerList.h:48-63, with the C alignment rules of x86-64


def write_sslist(path, seq_length, seed, packed, positions):
    """The on-disk sorted mer list `<fasta>.sslist` that DNAFileSML::Create leaves (LM/FileSML.cpp:401-459) and FileSML::LoadFile2
    reads back (:120-195): SMLHeader, the 2-bit sequence (ceil(2n/32) + 2 words), the uint32 positions in sorted order.  A list
    written here is picked up by the unmodified binary instead of being rebuilt (MatchList::LoadSMLs, LM/MatchList.h:296-330:
    format version and seed pattern must match).  Header fields the reference leaves uninitialised (word size, byte-order flag,
    description: they hold stray heap bytes in its files and are never read) are written as defined values."""
    packed = np.ascontiguousarray(packed, dtype=np.uint32)
    positions = np.ascontiguousarray(positions, dtype=np.uint32)
    L = getSeedLength(seed)
    words = (2 * seq_length) // 32 + (1 if (2 * seq_length) % 32 else 0) + 2
    if packed.size != words or positions.size != max(seq_length - L + 1, 0):
        raise ValueError("write_sslist: array sizes do not match a sequence of %d bases" % seq_length)
    h = np.zeros(SML_HEADER_BYTES, dtype=np.uint8)
    import struct
    struct.pack_into("<IIQIIQII", h, 0, SML_FORMAT_VERSION, 2, seed, L, getSeedWeight(seed), seq_length, 0xFFFFFFFF, 32)
    h[40] = 1  # little_endian; id (int16 at 42) and circular (44) stay 0
    table = np.zeros(255, dtype=np.uint8)  # SortedMerList::BasicDNATable, LM/SortedMerList.cpp:29-47
    for letters, code in (("cCbByY", 1), ("gGsSkK", 2), ("tT", 3)):
        for ch in letters:
            table[ord(ch)] = code
    h[45:300] = table
    with open(path, "wb") as f:
        f.write(h.tobytes())
        f.write(packed.tobytes())
        f.write(positions.tobytes())


def read_sslist(path):
    """FileSML::LoadFile2 (LM/FileSML.cpp:120-195) for a `<fasta>.sslist` file: returns (code, header, packed, positions) with the
    reference's return codes -- 0 ok, 1 file cannot be opened, 2 short header, 3 format version other than DNAFileSML's, 4 short
    sequence data, 5 position array shorter than header.length entries -- and None for what could not be read.  header: dict of the
    SMLHeader fields the reference reads (LM/SortedMerList.h:48-63).  packed: ceil(2n/32) + 2 words; positions: the uint32 array of
    n - seed_length + 1 sorted ranks (the file reserves `length` entries after the sequence, :175-176; DNAFileSML::Create writes
    SMLLength() of them, LM/FileSML.cpp:432-445)."""
    import struct
    try:
        f = open(path, "rb")
    except OSError:
        return 1, None, None, None
    with f:
        raw = f.read(SML_HEADER_BYTES)
        if len(raw) < SML_HEADER_BYTES:
            return 2, None, None, None
        version, abits, seed, slen, sweight, length, unique_mers, word_size = struct.unpack_from("<IIQIIQII", raw, 0)
        if version != SML_FORMAT_VERSION:
            return 3, None, None, None
        header = {"version": version, "alphabet_bits": abits, "seed": seed, "seed_length": slen, "seed_weight": sweight, "length": length,
                  "unique_mers": unique_mers, "word_size": word_size, "little_endian": raw[40], "circular": raw[44],
                  "translation_table": np.frombuffer(raw, dtype=np.uint8, count=255, offset=45).copy()}
        seq_len = length + (slen - 1 if header["circular"] else 0)
        words = (seq_len * abits) // 32 + (1 if (seq_len * abits) % 32 else 0) + 2
        buf = f.read(4 * words)
        if len(buf) < 4 * words:
            return 4, header, None, None
        packed = np.frombuffer(buf, dtype="<u4").copy()
        npos = max(length - slen + 1, 0) if slen else 0
        buf = f.read(4 * npos)
        if len(buf) < 4 * npos:
            return 5, header, packed, None
        return 0, header, packed, np.frombuffer(buf, dtype="<u4").copy()


def WriteList(rows, stream, seq_filenames=("null", "null"), seq_lengths=(0, 0)):
    """The match-list text format `progressiveMauve --mums` writes and `--match-input` reads back
    (WriteList, LM/MatchList.h:617-662; consumed by ReadList :526-614 at MA/progressiveMauve.cpp:472-491): a device-built
    list written here can be fed to the UNMODIFIED binary.  rows: [n, 3] (length, start0, start1) in list order, or a
    MatchList.  The reference prints each Match's address as its id (any distinct integer does: ReadList only keys a map with
    it); like the reference, nothing at all is written for an empty list."""
    if isinstance(rows, MatchList):
        rows = rows.as_array()
    rows = np.asarray(rows, dtype=np.int64).reshape(-1, 3)
    if rows.shape[0] == 0:
        return
    w = stream.write
    w("FormatVersion\t3\n")
    w("SequenceCount\t2\n")
    for i in range(2):
        w("Sequence%dFile\t%s\n" % (i, seq_filenames[i] if i < len(seq_filenames) else "null"))
        w("Sequence%dLength\t%d\n" % (i, seq_lengths[i] if i < len(seq_lengths) else 0))
    w("MatchCount\t%d\n" % rows.shape[0])
    for k, (ln, s0, s1) in enumerate(rows.tolist()):
        w("%d\t%d\t%d\t%d\t0\t0\n" % (ln, s0, s1, k + 1))


def ReadList(stream):
    """ReadList (LM/MatchList.h:526-614) for two sequences -> (rows[n,3], seq_filenames, seq_lengths); the same format errors raise ValueError"""
    tok = stream.read().split("\n")
    head = [l for l in tok[:7]]
    def field(line, tag):
        parts = line.split("\t", 1)
        if parts[0] != tag:
            raise ValueError("InvalidFileFormat: expected %s, found %r" % (tag, parts[0]))
        return parts[1] if len(parts) > 1 else ""
    if len(head) < 7 or field(head[0], "FormatVersion").strip() != "3":
        raise ValueError("InvalidFileFormat: FormatVersion 3 expected")
    if int(field(head[1], "SequenceCount")) != 2:
        raise ValueError("only two-sequence lists are handled here")
    names = [field(head[2], "Sequence0File"), field(head[4], "Sequence1File")]
    lens = [int(field(head[3], "Sequence0Length")), int(field(head[5], "Sequence1Length"))]
    count = int(field(head[6], "MatchCount"))
    rows = []
    for line in tok[7:]:
        if not line.strip():
            continue
        f = line.split()
        if int(f[4]) > 0:
            raise ValueError("Unable to read file, invalid format, cannot read subset data")
        rows.append((int(f[0]), int(f[1]), int(f[2])))
    if len(rows) != count:
        raise ValueError("InvalidFileFormat: MatchCount %d but %d rows" % (count, len(rows)))
    return np.array(rows, dtype=np.int64).reshape(-1, 3), names, lens


def _find_mums(seq0, seq1, seed, rule):
    a0, n0, k0 = _buf(seq0)
    a1, n1, k1 = _buf(seq1)
    out = C.POINTER(_capi.Match)()
    n_out = C.c_uint64(0)
    stats = np.zeros(8, dtype=np.uint64)
    check(lib().mcu_find_mums(a0, n0, a1, n1, seed, rule, C.byref(out), C.byref(n_out), stats.ctypes.data))
    n = n_out.value
    if n:
        rows = np.ctypeslib.as_array(C.cast(out, C.POINTER(C.c_int64)), shape=(n, 3)).copy()
    else:
        rows = np.zeros((0, 3), dtype=np.int64)
    lib().mcu_free(out)
    return rows, stats


class MemHash:
    """mems::MemHash for two genomes (repeat_tolerance 0, enumeration_tolerance 1: LM/MemHash.cpp:139-162)."""
    _rule = _capi.RULE_MEMHASH

    def __init__(self):
        self.Clear()

    def Clear(self):
        self.m_mem_count = 0
        self.m_collision_count = 0
        self.last_stats = np.zeros(8, dtype=np.uint64)
        self._rows = np.zeros((0, 3), dtype=np.int64)

    def MemCount(self):
        return self.m_mem_count

    def MemCollisionCount(self):
        return self.m_collision_count

    def FindMatches(self, ml: MatchList):
        """MemHash::FindMatches(MatchList&) (LM/MemHash.cpp:109-127): reads ml.seq_table / ml.sml_table, appends into ml."""
        if len(ml.seq_table) != 2:
            raise McuError(_capi.MCU_EINVAL, "the CUDA match finder handles exactly two genomes (PairwiseMatchFinder / gap_mh case)")
        if len(ml.sml_table) != 2:
            raise McuError(_capi.MCU_EINVAL, "MatchList has no sorted mer lists (call CreateMemorySMLs first)")
        seed = ml.sml_table[0].Seed()
        if ml.sml_table[1].Seed() != seed:
            raise McuError(_capi.MCU_EINVAL, "sorted mer lists were built with different seed patterns")
        rows, stats = _find_mums(ml.seq_table[0], ml.seq_table[1], seed, self._rule)
        self._rows, self.last_stats = rows, stats
        self.m_mem_count = int(stats[1])
        self.m_collision_count = int(stats[2])
        ml.matches.extend(Match(int(r[0]), [int(r[1]), int(r[2])]) for r in rows)
        return True

    def GetMatchList(self):
        return [Match(int(r[0]), [int(r[1]), int(r[2])]) for r in self._rows]

    def rows(self):
        return self._rows


class PairwiseMatchFinder(MemHash):
    """mems::PairwiseMatchFinder (LM/PairwiseMatchFinder.cpp:37-71), the finder progressiveMauve uses for <= 4 genomes."""
    _rule = _capi.RULE_PAIRWISE


def find_mums(seq0, seq1, seed, rule=_capi.RULE_PAIRWISE):
    """(rows[n,3] int64 in reference list order, stats[8]) straight from mcu_find_mums."""
    return _find_mums(seq0, seq1, seed, rule)


def find_mums_batch(pairs, seeds=None, rule=_capi.RULE_MEMHASH, seed_rank=0):
    """Gap search over many sequence pairs at once (mcu_find_mums_batch): the device form of one round of
    pairwiseAnchorSearch calls (LM/ProgressiveAligner.cpp:590-679).  pairs: [(bytes, bytes), ...]; seeds: one pattern per
    pair, default = the reference's choice getSeed(getDefaultSeedWeight((len0 + len1) / 2), rank 0) with 0 (no search) when
    the weight is below 5 (:617-627).  Returns ([rows[n_i,3] per pair, coordinates local to the pair], stats[4])."""
    n = len(pairs)
    if seeds is None:
        seeds = []
        for a, b in pairs:
            w = getDefaultSeedWeight((len(a) + len(b)) // 2)
            seeds.append(getSeed(w, seed_rank) if w >= 5 else 0)
    seeds = np.asarray(seeds, dtype=np.uint64)
    cat0 = b"".join(bytes(a) for a, _ in pairs)
    cat1 = b"".join(bytes(b) for _, b in pairs)
    off0 = np.zeros(n + 1, dtype=np.uint64)
    off1 = np.zeros(n + 1, dtype=np.uint64)
    off0[1:] = np.cumsum([len(a) for a, _ in pairs], dtype=np.uint64)
    off1[1:] = np.cumsum([len(b) for _, b in pairs], dtype=np.uint64)
    out = C.POINTER(_capi.Match)()
    out_off = np.zeros(n + 1, dtype=np.uint64)
    stats = np.zeros(4, dtype=np.uint64)
    check(lib().mcu_find_mums_batch(n, cat0, off0.ctypes.data, cat1, off1.ctypes.data, seeds.ctypes.data, rule, C.byref(out),
                                    out_off.ctypes.data, stats.ctypes.data))
    total = int(out_off[n])
    rows = np.ctypeslib.as_array(C.cast(out, C.POINTER(C.c_int64)), shape=(max(total, 1), 3))[:total].copy() if total else np.zeros((0, 3), dtype=np.int64)
    lib().mcu_free(out)
    return [rows[int(out_off[i]):int(out_off[i + 1])] for i in range(n)], stats


class AnchorSession:
    """Device-resident form of the same path (include/mauve_cuda.h: mcu_session_*): used by bench.py and multi-GPU runs."""

    def __init__(self):
        h = C.c_void_p()
        check(lib().mcu_session_create(C.byref(h)))
        self._h = h
        self.stage_ms = np.zeros(16, dtype=np.float32)
        self.stats = np.zeros(8, dtype=np.uint64)

    def close(self):
        if self._h:
            lib().mcu_session_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload(self, seq0, seq1):
        a0, n0, k0 = _buf(seq0)
        a1, n1, k1 = _buf(seq1)
        check(lib().mcu_session_upload(self._h, a0, n0, a1, n1))

    def upload_ptr(self, addr0, n0, addr1, n1):
        check(lib().mcu_session_upload(self._h, addr0, n0, addr1, n1))

    def upload_begin_ptr(self, addr0, n0, addr1, n1, chunks=8):
        """asynchronous chunked H2D (copy stream); the next run() overlaps pack + the first partition pass with the copies"""
        check(lib().mcu_session_upload_begin(self._h, addr0, n0, addr1, n1, chunks))

    def upload_sharded_ptr(self, addr0, n0, addr1, n1):
        """collective: this rank's 1 / world slice of both genomes (mcu_comm_init first)"""
        check(lib().mcu_session_upload_sharded(self._h, addr0, n0, addr1, n1))

    def run(self, seed, shard_index=0, shard_count=1):
        check(lib().mcu_session_run(self._h, seed, shard_index, shard_count, self.stage_ms.ctypes.data, self.stats.ctypes.data))
        return int(lib().mcu_session_match_count(self._h))

    def run_sharded(self, seed):
        """collective (every rank): one pass sharded over the ranks of mcu_comm_init; rank 0's session ends up with the whole list.
        Returns the total number of matches (on every rank)."""
        check(lib().mcu_session_run_sharded(self._h, seed, self.stage_ms.ctypes.data, self.stats.ctypes.data))
        return int(self.stats[1])

    def enumerate(self, seed, shard_index=0, shard_count=1):
        check(lib().mcu_session_enumerate(self._h, seed, shard_index, shard_count))

    def uniq_bitmap(self):
        """(device pointer, n_words) of the unique-seed bitmap: ranks SUM-all-reduce it between enumerate() and finish()"""
        p, n = C.c_void_p(), C.c_uint64(0)
        check(lib().mcu_session_uniq_bitmap(self._h, C.byref(p), C.byref(n)))
        return p.value, int(n.value)

    def finish(self, uniq_is_global=False):
        check(lib().mcu_session_finish(self._h, 1 if uniq_is_global else 0, self.stage_ms.ctypes.data, self.stats.ctypes.data))
        return int(lib().mcu_session_match_count(self._h))

    def merge(self, rows, in_device=False, n=None):
        """rank-0 merge of the gathered per-rank rows of a run finished with a global bitmap -> (match count, [replayed buckets, duplicate rows])"""
        if in_device:
            addr, cnt = rows, n
        else:
            rows = np.ascontiguousarray(rows, dtype=np.int64).reshape(-1, 3)
            addr, cnt = rows.ctypes.data, rows.shape[0]
        st = np.zeros(2, dtype=np.uint64)
        check(lib().mcu_session_merge(self._h, addr, cnt, 1 if in_device else 0, st.ctypes.data))
        return int(lib().mcu_session_match_count(self._h)), st

    def match_count(self):
        return int(lib().mcu_session_match_count(self._h))

    def download(self, out=None):
        n = self.match_count()
        if out is None:
            out = np.empty((n, 3), dtype=np.int64)
        check(lib().mcu_session_download(self._h, out.ctypes.data))
        return out[:n]

    def download_ptr(self, addr):
        check(lib().mcu_session_download(self._h, addr))

    def matches_device(self):
        return lib().mcu_session_matches_device(self._h)

    def launch_count(self):
        return int(lib().mcu_session_launch_count(self._h))


def merge_matches(rows, in_device=False, n=None, return_unclean=False):
    """Rank-0 merge of per-shard match lists (reference list order, duplicates dropped).  With return_unclean also
    returns the number of order-dependent hash buckets (non-zero -> re-run unsharded for an exact list)."""
    if in_device:
        addr, cnt = rows, n
    else:
        rows = np.ascontiguousarray(rows, dtype=np.int64).reshape(-1, 3)
        addr, cnt = rows.ctypes.data, rows.shape[0]
    out = C.POINTER(_capi.Match)()
    n_out = C.c_uint64(0)
    unclean = C.c_uint64(0)
    check(lib().mcu_merge_matches(addr, cnt, 1 if in_device else 0, C.byref(out), C.byref(n_out), C.byref(unclean)))
    k = n_out.value
    res = np.ctypeslib.as_array(C.cast(out, C.POINTER(C.c_int64)), shape=(k, 3)).copy() if k else np.zeros((0, 3), dtype=np.int64)
    lib().mcu_free(out)
    return (res, int(unclean.value)) if return_unclean else res


# ---- sorted mer list sharded by mer range -------------------------------------------------------------------------
def sml_build_shard(seq, seed, shard, n_shards, want_mers=True):
    """shard `shard` of `n_shards` of DNAMemorySML::Create's sorted list: (positions, mers) of the seeds whose canonical mer lies in the
    shard's range; the shards' lists, one after the other, are the whole list (equal mers in unspecified order inside their run)"""
    addr, n, keep = _buf(seq)
    L = getSeedLength(seed)
    m = max(n - L + 1, 1)
    pos = np.zeros(m, dtype=np.uint32)
    mer = np.zeros(m, dtype=np.uint64) if want_mers else None
    out_len = C.c_uint64(0)
    check(lib().mcu_sml_build_shard(addr, n, seed, int(shard), int(n_shards), pos.ctypes.data, mer.ctypes.data if want_mers else None, C.byref(out_len)))
    k = int(out_len.value)
    return (pos[:k].copy(), mer[:k].copy()) if want_mers else pos[:k].copy()


def sml_build_sharded(seq, seed):
    """collective over the communicator of dist.init_from_env(): every rank passes the same sequence; rank 0 gets the positions of the
    whole sorted list (others: None).  Returns (positions or None, device ms on this rank)."""
    addr, n, keep = _buf(seq)
    L = getSeedLength(seed)
    pos = np.zeros(max(n - L + 1, 1), dtype=np.uint32)
    out_len = C.c_uint64(0)
    ms = C.c_float(0)
    check(lib().mcu_sml_build_sharded(addr, n, seed, pos.ctypes.data, C.byref(out_len), C.byref(ms)))
    return pos[:int(out_len.value)], float(ms.value)


# ---- LM/ProgressiveAligner.h:300-406, LM/MatchList.h:680-692, LM/GreedyBreakpointElimination.h:161-250 -------------
def EliminateOverlaps_v2(rows, eliminate_both=False, min_length=0, return_ties=False):
    """mems::EliminateOverlaps_v2(ml, eliminate_both) on a two-genome match list (rows: len, start0 > 0, start1 signed), followed by
    ml.LengthFilter(min_length) when min_length > 0 -> the rows the reference's list holds afterwards, in its order."""
    rows = np.ascontiguousarray(rows, dtype=np.int64).reshape(-1, 3)
    out = np.zeros_like(rows)
    n_out, ties = C.c_uint64(0), C.c_uint64(0)
    check(lib().mcu_eliminate_overlaps(rows.ctypes.data, rows.shape[0], 1 if eliminate_both else 0, int(min_length), out.ctypes.data, C.byref(n_out),
                                       C.byref(ties)))
    res = out[:int(n_out.value)].copy()
    return (res, int(ties.value)) if return_ties else res


def IdentifyBreakpoints(rows, return_ties=False):
    """mems::IdentifyBreakpoints -> (the list ordered on genome 0, breakpoints = index of the last match of every LCB)"""
    rows = np.ascontiguousarray(rows, dtype=np.int64).reshape(-1, 3)
    out = np.zeros_like(rows)
    bp = np.zeros(rows.shape[0] + 1, dtype=np.uint64)
    n_bp, ties = C.c_uint64(0), C.c_uint64(0)
    check(lib().mcu_lcbs(rows.ctypes.data, rows.shape[0], out.ctypes.data, bp.ctypes.data, C.byref(n_bp), C.byref(ties)))
    res = (out, bp[:int(n_bp.value)].copy())
    return res + (int(ties.value),) if return_ties else res


def ComputeLCBs_v2(sorted_rows, breakpoints):
    """mems::ComputeLCBs_v2: the LCBs as a list of row blocks (views of sorted_rows)"""
    out, prev = [], 0
    for b in np.asarray(breakpoints, dtype=np.int64).tolist():
        out.append(sorted_rows[prev:b + 1])
        prev = b + 1
    return out


# ---- MU/anchoredpp.cpp:256-409, MU/anchors.cpp:9-186 ----------------------------------------------------------------
class AnchorParams(C.Structure):
    """mcu_anchor_params (include/mauve_cuda.h): the MUSCLE globals the column scoring reads.  AnchorParams.default() = the DNA
    settings MuscleInterface::ProfileAlignFast leaves in force (LM/MuscleInterface.cpp:1086-1106)."""
    _fields_ = [("subst", C.c_float * 16), ("gap_open", C.c_float), ("gap_extend", C.c_float), ("term_gap", C.c_float),
                ("smooth_ceil", C.c_float), ("min_best_col", C.c_float), ("min_smooth", C.c_float),
                ("smooth_window", C.c_uint32), ("anchor_spacing", C.c_uint32), ("letter_of_char", C.c_uint8 * 256)]

    @classmethod
    def default(cls):
        p = cls()
        lib().mcu_anchor_default_params(C.addressof(p))
        return p


def FindAnchorColsPP_batch(windows, params=None, return_scores=False):
    """muscle::FindAnchorColsPP for many windows in one device call.  windows: sequence of (rows, n1[, weights]) with rows =
    uint8[(n1 + n2), ncol] characters ('-' / '.' gaps; the first alignment's n1 rows first) and weights = MSA::GetSeqWeight per row
    (None: all 1).  -> list of uint32 arrays of anchor columns (with return_scores: (cols, MatchScore, SmoothScore) per window)"""
    n = len(windows)
    if n == 0:
        return []
    mats, n1s, n2s, ws, any_w = [], [], [], [], False
    for win in windows:
        rows, n1 = win[0], int(win[1])
        w = win[2] if len(win) > 2 else None
        rows = np.ascontiguousarray(rows, dtype=np.uint8)
        if rows.ndim != 2 or not (0 < n1 < rows.shape[0]):
            raise ValueError("a window is uint8[(n1 + n2), ncol] with both alignments non-empty")
        mats.append(rows)
        n1s.append(n1)
        n2s.append(rows.shape[0] - n1)
        any_w = any_w or w is not None
        ws.append(np.ones(rows.shape[0], dtype=np.float32) if w is None else np.ascontiguousarray(w, dtype=np.float32))
        if ws[-1].size != rows.shape[0]:
            raise ValueError("one weight per row")
    ncol = np.array([m.shape[1] for m in mats], dtype=np.uint32)
    row_off = np.zeros(n + 1, dtype=np.uint64)
    row_off[1:] = np.cumsum([m.size for m in mats])
    col_off = np.zeros(n + 1, dtype=np.uint64)
    col_off[1:] = np.cumsum(ncol.astype(np.uint64))
    blob = np.concatenate([m.reshape(-1) for m in mats]) if int(row_off[-1]) else np.zeros(1, dtype=np.uint8)
    weights = np.concatenate(ws) if any_w else None
    total = max(int(col_off[-1]), 1)
    cols = np.zeros(total, dtype=np.uint32)
    counts = np.zeros(n, dtype=np.uint32)
    score = np.zeros(total, dtype=np.float32) if return_scores else None
    smooth = np.zeros(total, dtype=np.float32) if return_scores else None
    n1a, n2a = np.array(n1s, dtype=np.uint32), np.array(n2s, dtype=np.uint32)
    check(lib().mcu_anchor_cols_batch(n, blob.ctypes.data, row_off.ctypes.data, ncol.ctypes.data, n1a.ctypes.data, n2a.ctypes.data,
                                      weights.ctypes.data if weights is not None else None, C.addressof(params) if params is not None else None,
                                      col_off.ctypes.data, cols.ctypes.data, counts.ctypes.data,
                                      score.ctypes.data if return_scores else None, smooth.ctypes.data if return_scores else None, None))
    out = []
    for i in range(n):
        a, b = int(col_off[i]), int(col_off[i + 1])
        c = cols[a:a + int(counts[i])].copy()
        out.append((c, score[a:b].copy(), smooth[a:b].copy()) if return_scores else c)
    return out


def FindAnchorColsPP(msa1, msa2, weights=None, params=None, return_scores=False):
    """muscle::FindAnchorColsPP(msa1, msa2, AnchorCols, &count): msa1 / msa2 = the two alignments of a window as uint8[rows, ncol]
    (or one bytes row each) -> the anchor columns.  Alignments of different lengths have none (MU/anchoredpp.cpp:358-362)."""
    a = np.atleast_2d(np.frombuffer(msa1, dtype=np.uint8) if isinstance(msa1, (bytes, bytearray)) else np.asarray(msa1, dtype=np.uint8))
    b = np.atleast_2d(np.frombuffer(msa2, dtype=np.uint8) if isinstance(msa2, (bytes, bytearray)) else np.asarray(msa2, dtype=np.uint8))
    if a.shape[1] != b.shape[1]:
        z = np.zeros(0, dtype=np.uint32)
        return (z, np.zeros(0, np.float32), np.zeros(0, np.float32)) if return_scores else z
    return FindAnchorColsPP_batch([(np.concatenate([a, b]), a.shape[0], weights)], params=params, return_scores=return_scores)[0]


# ---- MU/pwpath.h, MU/glbalign.cpp ------------------------------------------------------------------
@dataclass
class PWPath:
    """Edge types of a pairwise path, first edge first: 'M' both advance, 'D' only A, 'I' only B (MU/pwpath.h:46-51)."""
    edges: bytes
    score: int

    def GetEdgeCount(self):
        return len(self.edges)


def GlobalAlignBatch(pairs, return_ms=False):
    """muscle::GlobalAlign for many (a, b) ACGT string pairs at once; returns one PWPath per pair."""
    n = len(pairs)
    if n == 0:
        return ([], 0.0) if return_ms else []
    a_off = np.zeros(n + 1, dtype=np.uint64)
    b_off = np.zeros(n + 1, dtype=np.uint64)
    la = np.fromiter((len(p[0]) for p in pairs), dtype=np.uint64, count=n)
    lb = np.fromiter((len(p[1]) for p in pairs), dtype=np.uint64, count=n)
    np.cumsum(la, out=a_off[1:])
    np.cumsum(lb, out=b_off[1:])
    a = np.frombuffer(b"".join(bytes(p[0]) for p in pairs), dtype=np.uint8)
    b = np.frombuffer(b"".join(bytes(p[1]) for p in pairs), dtype=np.uint8)
    res = nw_batch_arrays(a, a_off, b, b_off)
    paths = [PWPath(res["path"][int(res["path_off"][i]):int(res["path_off"][i]) + int(res["path_len"][i])].tobytes(), int(res["score"][i]))
             for i in range(n)]
    return (paths, res["device_ms"]) if return_ms else paths


def nw_batch_arrays(a, a_off, b, b_off):
    """Array form of mcu_nw_batch: concatenated sequences + offsets in, path buffer + lengths + scores out."""
    a = np.ascontiguousarray(a, dtype=np.uint8)
    b = np.ascontiguousarray(b, dtype=np.uint8)
    a_off = np.ascontiguousarray(a_off, dtype=np.uint64)
    b_off = np.ascontiguousarray(b_off, dtype=np.uint64)
    n = a_off.size - 1
    path_off = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum((a_off[1:] - a_off[:-1]) + (b_off[1:] - b_off[:-1]), out=path_off[1:])
    path = np.empty(max(int(path_off[-1]), 1), dtype=np.uint8)
    path_len = np.zeros(max(n, 1), dtype=np.uint32)
    score = np.zeros(max(n, 1), dtype=np.int64)
    ms = C.c_float(0)
    check(lib().mcu_nw_batch(n, a.ctypes.data, a_off.ctypes.data, b.ctypes.data, b_off.ctypes.data, path_off.ctypes.data,
                             path.ctypes.data, path_len.ctypes.data, score.ctypes.data, C.byref(ms)))
    stats = np.zeros(5, dtype=np.uint64)
    lib().mcu_nw_last_stats(stats.ctypes.data)
    return {"path": path, "path_off": path_off, "path_len": path_len[:n], "score": score[:n], "device_ms": float(ms.value), "stats": stats}


def GlobalAlignBatchWild(pairs):
    """muscle::GlobalAlign for (a, b) pairs whose sequences may contain DNA wildcards (mcu_nw_batch_wild: the reference's float
    arithmetic, one thread per region); PWPath.score is the reference's float score."""
    n = len(pairs)
    if n == 0:
        return []
    a_off = np.zeros(n + 1, dtype=np.uint64)
    b_off = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum(np.fromiter((len(p[0]) for p in pairs), dtype=np.uint64, count=n), out=a_off[1:])
    np.cumsum(np.fromiter((len(p[1]) for p in pairs), dtype=np.uint64, count=n), out=b_off[1:])
    a = np.frombuffer(b"".join(bytes(p[0]) for p in pairs) or b"\0", dtype=np.uint8)
    b = np.frombuffer(b"".join(bytes(p[1]) for p in pairs) or b"\0", dtype=np.uint8)
    path_off = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum((a_off[1:] - a_off[:-1]) + (b_off[1:] - b_off[:-1]), out=path_off[1:])
    path = np.zeros(max(int(path_off[-1]), 1), dtype=np.uint8)
    path_len = np.zeros(n, dtype=np.uint32)
    score = np.zeros(n, dtype=np.float32)
    ms = C.c_float(0)
    check(lib().mcu_nw_batch_wild(n, a.ctypes.data, a_off.ctypes.data, b.ctypes.data, b_off.ctypes.data, path_off.ctypes.data,
                                  path.ctypes.data, path_len.ctypes.data, score.ctypes.data, C.byref(ms)))
    return [PWPath(path[int(path_off[i]):int(path_off[i]) + int(path_len[i])].tobytes(), float(score[i])) for i in range(n)]


def GlobalAlign(a, b) -> PWPath:
    return GlobalAlignBatch([(a, b)])[0]


# ---- LM/HomologyHMM -----------------------------------------------------------------------------------
@dataclass
class Params:
    """struct Params (LM/HomologyHMM/homology.h:169-177)"""
    iStartHomologous: float = 0.5
    iGoHomologous: float = 0.00001
    iGoUnrelated: float = 0.0000001
    iGoStopFromUnrelated: float = 0.0000001
    iGoStopFromHomologous: float = 0.0000001
    aEmitHomologous: List[float] = field(default_factory=lambda: [0.0] * 8)
    aEmitUnrelated: List[float] = field(default_factory=lambda: [0.0] * 8)

    def as_array(self):
        return np.array([self.iStartHomologous, self.iGoHomologous, self.iGoUnrelated, self.iGoStopFromUnrelated,
                         self.iGoStopFromHomologous] + list(self.aEmitHomologous) + list(self.aEmitUnrelated), dtype=np.float64)

    @staticmethod
    def from_array(v):
        v = [float(x) for x in v]
        return Params(v[0], v[1], v[2], v[3], v[4], v[5:13], v[13:21])


def getAdaptedHoxdMatrixParameters(gc_content: float) -> Params:
    out = np.zeros(21, dtype=np.float64)
    check(lib().mcu_hmm_params(gc_content, 0.0, 0.0, 0.0, out.ctypes.data))
    return Params.from_array(out)


def hmm_params(gc_content=0.5, go_homologous=0.0, go_unrelated=0.0, pct_identity=0.0):
    out = np.zeros(21, dtype=np.float64)
    check(lib().mcu_hmm_params(gc_content, go_homologous, go_unrelated, pct_identity, out.ctypes.data))
    return out


def adaptToPercentIdentity(params: Params, pct_identity: float) -> Params:
    """LM/HomologyHMM/parameters.h:140-159 (host arithmetic on 8 numbers; mutates and returns params)."""
    if pct_identity <= 0 or pct_identity > 1:
        raise ValueError("Bad pct identity")
    e = params.aEmitHomologous
    gapnorm = pct_identity * (1.0 - e[6] - e[7])
    prev = e[0] + e[1]
    diff = prev - gapnorm
    rest = e[2] + e[3] + e[4] + e[5]
    for i in (2, 3, 4, 5):
        e[i] += diff * e[i] / rest
    for i in (0, 1):
        e[i] -= diff * e[i] / prev
    return params


def run_batch(sequences, params, want_posterior=False):
    """run() for many symbol strings ('1'..'8') at once -> list of 'H'/'N' predictions (+ posteriors)."""
    p = params.as_array() if isinstance(params, Params) else np.ascontiguousarray(params, dtype=np.float64)
    n = len(sequences)
    off = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum(np.fromiter((len(s) for s in sequences), dtype=np.uint64, count=n), out=off[1:])
    sym = np.frombuffer(b"".join(bytes(s) for s in sequences), dtype=np.uint8)
    total = int(off[-1])
    pred = np.empty(max(total, 1), dtype=np.uint8)
    post = np.empty(max(total, 1), dtype=np.float64) if want_posterior else None
    ms = C.c_float(0)
    symaddr = sym.ctypes.data if total else np.zeros(1, dtype=np.uint8).ctypes.data
    check(lib().mcu_hmm_batch(n, symaddr, off.ctypes.data, p.ctypes.data, pred.ctypes.data,
                              post.ctypes.data if want_posterior else None, C.byref(ms)))
    preds = [pred[int(off[i]):int(off[i + 1])].tobytes() for i in range(n)]
    if want_posterior:
        return preds, [post[int(off[i]):int(off[i + 1])] for i in range(n)], float(ms.value)
    return preds


def run(sequence, params, want_posterior=False):
    """void run(std::string& sequence, std::string& prediction, const Params&) (LM/HomologyHMM/homology.h:47)"""
    if want_posterior:
        preds, posts, _ = run_batch([sequence], params, True)
        return preds[0], posts[0]
    return run_batch([sequence], params)[0]


# ---- test hook --------------------------------------------------------------------------------------
# ---- LM/SeedOccurrenceList.h, LM/GreedyBreakpointElimination.h:403-476 -----------------------------------
hoxd_matrix = np.array([[91, -114, -31, -123], [-114, 100, -125, -31], [-31, -125, 100, -114], [-123, -31, -114, 91]],
                       dtype=np.int32)  # LM/SubstitutionMatrix.h:23-33


class SeedOccurrenceList:
    """mems::SeedOccurrenceList (LM/SeedOccurrenceList.h): construct(sml) computes the smoothed seed multiplicity of every
    position of the sml's sequence on the device; getFrequency(position) reads it."""

    def __init__(self):
        self._freq = np.zeros(0, dtype=np.float32)

    def construct(self, sml: DNAMemorySML):
        seq = sml._seq
        a, n, keep = _buf(seq)
        out = np.zeros(max(n, 1), dtype=np.float32)
        check(lib().mcu_sol_build(a, n, sml.Seed(), out.ctypes.data))
        self._freq = out[:n]

    def getFrequency(self, position: int) -> float:
        return float(self._freq[position])

    def frequencies(self):
        return self._freq


def anchor_scores(seq0, seq1, rows, lcb_off, seed=0, sol_1=None, sol_2=None, matrix=None, penalize_repeats=False):
    """mcu_anchor_scores: (lcb_scores float64[n_lcb], match_scores int64[n_rows]).  rows: int64 [n,3] (len, start0, start1);
    lcb_off: n_lcb+1 row offsets; sol_1/sol_2: SeedOccurrenceList (None = built on the device from `seed`)."""
    a0, n0, k0 = _buf(seq0)
    a1, n1, k1 = _buf(seq1)
    rows = np.ascontiguousarray(rows, dtype=np.int64).reshape(-1, 3)
    off = np.ascontiguousarray(lcb_off, dtype=np.uint64)
    n_lcb = max(off.size - 1, 0)
    f0 = None if sol_1 is None else np.ascontiguousarray(sol_1.frequencies(), dtype=np.float32)
    f1 = None if sol_2 is None else np.ascontiguousarray(sol_2.frequencies(), dtype=np.float32)
    for f, n in ((f0, n0), (f1, n1)):
        if f is not None and f.size != n:
            raise McuError(_capi.MCU_EINVAL, "seed occurrence list length differs from the sequence length")
    mat = None if matrix is None else np.ascontiguousarray(matrix, dtype=np.int32).reshape(16)
    lcb = np.zeros(max(n_lcb, 1), dtype=np.float64)
    ms = np.zeros(max(rows.shape[0], 1), dtype=np.int64)
    check(lib().mcu_anchor_scores(a0, n0, a1, n1, seed, None if f0 is None else f0.ctypes.data, None if f1 is None else f1.ctypes.data,
                                  rows.ctypes.data, rows.shape[0], off.ctypes.data, n_lcb, None if mat is None else mat.ctypes.data,
                                  1 if penalize_repeats else 0, lcb.ctypes.data, ms.ctypes.data))
    return lcb[:n_lcb], ms[:rows.shape[0]]


def GetPairwiseAnchorScore(lcb, seq_table, subst_scoring, sol_1, sol_2, penalize_gaps=False, penalize_repeats=False):
    """mems::GetPairwiseAnchorScore (LM/GreedyBreakpointElimination.h:403-476) for one LCB: a list of Match (ungapped, two
    genomes).  subst_scoring: 4x4 matrix or None for the default scheme."""
    if penalize_gaps:
        raise McuError(_capi.MCU_EINVAL, "penalize_gaps: ungapped matches have no gaps to penalize; the aligner never sets it")
    rows = np.array([[m.Length(), m.Start(0), m.Start(1)] for m in lcb], dtype=np.int64).reshape(-1, 3)
    scores, _ = anchor_scores(seq_table[0], seq_table[1], rows, [0, rows.shape[0]], 0, sol_1, sol_2, subst_scoring, penalize_repeats)
    return float(scores[0])


def sort_pairs(keys, vals, bits):
    keys = np.ascontiguousarray(keys).copy()
    vals = np.ascontiguousarray(vals, dtype=np.uint32).copy()
    assert keys.dtype in (np.uint32, np.uint64)
    check(lib().mcu_test_sort_pairs(keys.ctypes.data, vals.ctypes.data, keys.size, keys.dtype.itemsize, bits))
    return keys, vals
