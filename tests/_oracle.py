"""ctypes bindings for the CPU checkers (TEST INFRASTRUCTURE ONLY).

`oracle()`  -> oracle/libmauve_oracle.so : our C restatement (oracle/mauve_oracle.c)
`ref()`     -> oracle/_ref/libmauve_ref.so : the reference's own sources compiled in place
               (only present where oracle/Makefile.ref was run; tests skip when absent).
Nothing under mauve_py_b200/ imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "libmauve_oracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libmauve_ref.so")
REF_FULL_SO = os.path.join(ROOT, "oracle", "_ref", "libmauve_ref_full.so")  # all of libMems: the rows next to the hot path


class Match3(C.Structure):
    _fields_ = [("len", C.c_int64), ("s0", C.c_int64), ("s1", C.c_int64)]


_cache = {}


def _proto(lib, prefix):
    u64, p = C.c_uint64, C.c_void_p
    g = lambda n: getattr(lib, prefix + n)
    g("get_seed").restype = u64
    g("get_seed").argtypes = [C.c_int, C.c_int]
    g("default_seed_weight").restype = C.c_uint
    g("default_seed_weight").argtypes = [u64]
    g("seed_length").argtypes = [u64]
    g("seed_weight").argtypes = [u64]
    g("sml_build").restype = C.c_longlong
    g("sml_build").argtypes = [C.c_char_p, u64, u64, p, p]
    g("find_mums").restype = C.c_longlong
    g("find_mums").argtypes = [C.c_char_p, u64, C.c_char_p, u64, u64, C.c_int, C.POINTER(C.POINTER(Match3)), p]
    g("free").argtypes = [p]
    g("nw_align").restype = C.c_longlong
    g("hmm_params").argtypes = [C.c_double] * 4 + [p]
    g("hmm_run").argtypes = [C.c_char_p, u64, p, p, p]


def oracle(build=True):
    if "o" not in _cache:
        if build and (not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(
                os.path.join(ROOT, "oracle", "mauve_oracle.c"))):
            subprocess.check_call(["make", "-s", "-f", os.path.join(ROOT, "oracle", "Makefile")])
        lib = C.CDLL(ORACLE_SO)
        _proto(lib, "orc_")
        lib.orc_nw_align.argtypes = [C.c_char_p, C.c_uint, C.c_char_p, C.c_uint, C.c_void_p, C.c_void_p]
        lib.orc_hmm_encode.restype = C.c_longlong
        lib.orc_hmm_encode.argtypes = [C.c_char_p, C.c_char_p, C.c_uint64, C.c_void_p]
        lib.orc_pack.argtypes = [C.c_char_p, C.c_uint64, C.c_void_p]
        lib.orc_packed_words.restype = C.c_uint64
        lib.orc_packed_words.argtypes = [C.c_uint64]
        lib.orc_nw_align_f.restype = C.c_longlong
        lib.orc_nw_align_f.argtypes = [C.c_char_p, C.c_uint, C.c_char_p, C.c_uint, C.c_void_p, C.c_void_p]
        lib.orc_sol_build.restype = C.c_longlong
        lib.orc_sol_build.argtypes = [C.c_char_p, C.c_uint64, C.c_uint64, C.c_void_p]
        lib.orc_anchor_scores.argtypes = [C.c_char_p, C.c_uint64, C.c_char_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64,
                                          C.c_void_p, C.c_uint64, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        _cache["o"] = lib
    return _cache["o"]


def have_ref():
    return os.path.exists(REF_SO)


def ref():
    if "r" not in _cache:
        lib = C.CDLL(REF_SO)
        _proto(lib, "ref_")
        lib.ref_nw_align.argtypes = [C.c_char_p, C.c_uint, C.c_char_p, C.c_uint, C.c_void_p]
        _cache["r"] = lib
    return _cache["r"]


def nw_align_f(a: bytes, b: bytes):
    """NWSmall for sequences with DNA wildcards, in the reference's float arithmetic (orc_nw_align_f): (path bytes, float score)"""
    buf = np.zeros(len(a) + len(b) + 1, dtype=np.uint8)
    score = C.c_float(0)
    n = oracle().orc_nw_align_f(a, len(a), b, len(b), buf.ctypes.data, C.byref(score))
    if n < 0:
        raise RuntimeError("nw_align_f failed")
    return buf[:n].tobytes(), float(score.value)


def have_ref_full():
    return os.path.exists(REF_FULL_SO)


def ref_full():
    if "rf" not in _cache:
        lib = C.CDLL(REF_FULL_SO)
        u64, p = C.c_uint64, C.c_void_p
        lib.ref_sol_build.restype = C.c_longlong
        lib.ref_sol_build.argtypes = [C.c_char_p, u64, u64, p]
        lib.ref_anchor_scores.argtypes = [C.c_char_p, u64, C.c_char_p, u64, u64, p, u64, p, u64, C.c_int, p]
        lib.ref_eliminate_overlaps.restype = C.c_longlong
        lib.ref_eliminate_overlaps.argtypes = [p, u64, C.c_int, u64, p]
        lib.ref_lcbs.restype = C.c_longlong
        lib.ref_lcbs.argtypes = [p, u64, p, p]
        lib.ref_anchor_cols.restype = C.c_longlong
        lib.ref_anchor_cols.argtypes = [p, C.c_uint, C.c_uint, C.c_uint, p, C.c_int, p, p, p, p]
        lib.ref_anchor_settings.restype = None
        lib.ref_anchor_settings.argtypes = [p, p]
        _cache["rf"] = lib
    return _cache["rf"]


def eliminate_overlaps(rows, eliminate_both=False, min_length=0, use_ref=False):
    """EliminateOverlaps_v2 (+ LengthFilter) on a two-genome match list -> (rows afterwards in the reference's order, ties or None)"""
    rows = np.ascontiguousarray(rows, dtype=np.int64).reshape(-1, 3)
    out = np.zeros_like(rows)
    if use_ref:
        n = ref_full().ref_eliminate_overlaps(rows.ctypes.data, rows.shape[0], int(eliminate_both), int(min_length), out.ctypes.data)
        ties = None
    else:
        lib = oracle()
        lib.orc_eliminate_overlaps.restype = C.c_longlong
        lib.orc_eliminate_overlaps.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_uint64, C.c_void_p, C.c_void_p]
        t = C.c_uint64(0)
        n = lib.orc_eliminate_overlaps(rows.ctypes.data, rows.shape[0], int(eliminate_both), int(min_length), out.ctypes.data, C.byref(t))
        ties = int(t.value)
    if n < 0:
        raise RuntimeError("eliminate_overlaps failed")
    return out[:n].copy(), ties


def lcbs(rows, use_ref=False):
    """IdentifyBreakpoints + ComputeLCBs_v2 -> (rows sorted on genome 0, breakpoints = index of the last match of every LCB, ties or None)"""
    rows = np.ascontiguousarray(rows, dtype=np.int64).reshape(-1, 3)
    out = np.zeros_like(rows)
    bp = np.zeros(rows.shape[0] + 1, dtype=np.uint64)
    if use_ref:
        n = ref_full().ref_lcbs(rows.ctypes.data, rows.shape[0], out.ctypes.data, bp.ctypes.data)
        ties = None
    else:
        lib = oracle()
        lib.orc_lcbs.restype = C.c_longlong
        lib.orc_lcbs.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
        t = C.c_uint64(0)
        n = lib.orc_lcbs(rows.ctypes.data, rows.shape[0], out.ctypes.data, bp.ctypes.data, C.byref(t))
        ties = int(t.value)
    if n < 0:
        raise RuntimeError("lcbs failed")
    return out, bp[:n].copy(), ties


class AnchorParams(C.Structure):
    """orc_anchor_params == mcu_anchor_params (include/mauve_cuda.h)"""
    _fields_ = [("subst", C.c_float * 16), ("gap_open", C.c_float), ("gap_extend", C.c_float), ("term_gap", C.c_float),
                ("smooth_ceil", C.c_float), ("min_best_col", C.c_float), ("min_smooth", C.c_float),
                ("smooth_window", C.c_uint), ("anchor_spacing", C.c_uint), ("letter_of_char", C.c_uint8 * 256)]


def anchor_default_params():
    p = AnchorParams()
    lib = oracle()
    lib.orc_anchor_default_params.restype = None
    lib.orc_anchor_default_params.argtypes = [C.c_void_p]
    lib.orc_anchor_default_params(C.byref(p))
    return p


def anchor_settings_ref():
    """the settings the reference's column scoring reads after MuscleInterface's set-up -> (25 floats, 256 letters)"""
    out = np.zeros(25, dtype=np.float32)
    letters = np.zeros(256, dtype=np.uint8)
    ref_full().ref_anchor_settings(out.ctypes.data, letters.ctypes.data)
    return out, letters


def anchor_cols(rows, n1, weights=None, use_ref=False, params=None):
    """FindAnchorColsPP on a window: rows = uint8[(n1 + n2), ncol] characters ('-' gaps), the first alignment's n1 rows first.
    -> (anchor columns uint32, per-column score float32, smoothed score float32, weights float32, rows as scored).
    use_ref: the reference's own functions; weights None there = PrepareMSAforScoring computes them (as AnchoredProfileProfile does)"""
    rows = np.ascontiguousarray(rows, dtype=np.uint8)
    nr, ncol = rows.shape
    n2 = nr - n1
    cols = np.zeros(ncol + 1, dtype=np.uint32)
    score = np.zeros(ncol + 1, dtype=np.float32)
    smooth = np.zeros(ncol + 1, dtype=np.float32)
    if use_ref:
        w = np.ones(nr, dtype=np.float32) if weights is None else np.ascontiguousarray(weights, dtype=np.float32).copy()
        fixed = np.zeros_like(rows)
        n = ref_full().ref_anchor_cols(rows.ctypes.data, n1, n2, ncol, w.ctypes.data, 1 if weights is None else 0, cols.ctypes.data,
                                       score.ctypes.data, smooth.ctypes.data, fixed.ctypes.data)
        rows = fixed
    else:
        w = np.ones(nr, dtype=np.float32) if weights is None else np.ascontiguousarray(weights, dtype=np.float32)
        lib = oracle()
        lib.orc_anchor_cols.restype = C.c_longlong
        lib.orc_anchor_cols.argtypes = [C.c_void_p, C.c_uint, C.c_uint, C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        pr = params if params is not None else anchor_default_params()
        n = lib.orc_anchor_cols(rows.ctypes.data, n1, n2, ncol, w.ctypes.data, C.byref(pr), cols.ctypes.data, score.ctypes.data, smooth.ctypes.data)
    if n < 0:
        raise RuntimeError("anchor_cols failed")
    return cols[:n].copy(), score[:ncol].copy(), smooth[:ncol].copy(), w, rows


def sol_build(seq: bytes, seed: int, use_ref=False):
    """SeedOccurrenceList::construct: float32[n].  use_ref: the reference's own class (oracle/_ref), else the C restatement"""
    out = np.zeros(max(len(seq), 1), dtype=np.float32)
    f = ref_full().ref_sol_build if use_ref else oracle().orc_sol_build
    if f(seq, len(seq), seed, out.ctypes.data) != len(seq):
        raise RuntimeError("sol_build failed")
    return out[:len(seq)]


def anchor_scores(s0: bytes, s1: bytes, seed: int, rows, lcb_off, penalize_repeats=False, use_ref=False, freq=None, matrix=None):
    """GetPairwiseAnchorScore per LCB: (lcb_scores float64, match_scores int64 or None for the reference)"""
    rows = np.ascontiguousarray(rows, dtype=np.int64).reshape(-1, 3)
    off = np.ascontiguousarray(lcb_off, dtype=np.uint64)
    n_lcb = off.size - 1
    lcb = np.zeros(max(n_lcb, 1), dtype=np.float64)
    if use_ref:
        assert matrix is None
        if ref_full().ref_anchor_scores(s0, len(s0), s1, len(s1), seed, rows.ctypes.data, rows.shape[0], off.ctypes.data, n_lcb,
                                        int(penalize_repeats), lcb.ctypes.data) != 0:
            raise RuntimeError("ref_anchor_scores failed")
        return lcb[:n_lcb], None
    f0, f1 = freq if freq is not None else (sol_build(s0, seed), sol_build(s1, seed))
    ms = np.zeros(max(rows.shape[0], 1), dtype=np.int64)
    mat = None if matrix is None else np.ascontiguousarray(matrix, dtype=np.int32)
    if oracle().orc_anchor_scores(s0, len(s0), s1, len(s1), f0.ctypes.data, f1.ctypes.data, rows.ctypes.data, rows.shape[0], off.ctypes.data,
                                  n_lcb, None if mat is None else mat.ctypes.data, int(penalize_repeats), lcb.ctypes.data, ms.ctypes.data) != 0:
        raise RuntimeError("orc_anchor_scores failed")
    return lcb[:n_lcb], ms[:rows.shape[0]]


class Checker:
    """Uniform python face over either library (prefix 'orc_' or 'ref_')."""

    def __init__(self, lib, prefix):
        self.lib, self.prefix = lib, prefix

    def _f(self, name):
        return getattr(self.lib, self.prefix + name)

    def get_seed(self, weight, rank):
        return int(self._f("get_seed")(weight, rank))

    def default_seed_weight(self, avg_len):
        return int(self._f("default_seed_weight")(avg_len))

    def seed_length(self, seed):
        return int(self._f("seed_length")(seed))

    def seed_weight(self, seed):
        return int(self._f("seed_weight")(seed))

    def sml_build(self, seq: bytes, seed: int):
        n = len(seq)
        L = self.seed_length(seed)
        m = max(n - L + 1, 0)
        pos = np.zeros(max(m, 1), dtype=np.uint32)
        mer = np.zeros(max(m, 1), dtype=np.uint64)
        r = self._f("sml_build")(seq, n, seed, pos.ctypes.data, mer.ctypes.data)
        if r < 0:
            raise RuntimeError("sml_build failed")
        return pos[:r], mer[:r]

    def find_mums(self, s0: bytes, s1: bytes, seed: int, rule: int = 0):
        out = C.POINTER(Match3)()
        stats = np.zeros(4, dtype=np.uint64)
        n = self._f("find_mums")(s0, len(s0), s1, len(s1), seed, rule, C.byref(out), stats.ctypes.data)
        if n < 0:
            raise RuntimeError("find_mums failed")
        if n:
            arr = np.ctypeslib.as_array(C.cast(out, C.POINTER(C.c_int64)), shape=(n, 3)).copy()
        else:
            arr = np.zeros((0, 3), dtype=np.int64)
        self._f("free")(out)
        return arr, stats

    def nw_align(self, a: bytes, b: bytes):
        buf = np.zeros(len(a) + len(b) + 1, dtype=np.uint8)
        if self.prefix == "orc_":
            score = C.c_int64(0)
            n = self.lib.orc_nw_align(a, len(a), b, len(b), buf.ctypes.data, C.byref(score))
            sc = score.value
        else:
            n = self.lib.ref_nw_align(a, len(a), b, len(b), buf.ctypes.data)
            sc = None
        if n < 0:
            raise RuntimeError("nw_align failed")
        return buf[:n].tobytes(), sc

    def hmm_params(self, gc=0.5, go_h=0.0, go_u=0.0, pct_id=0.0):
        out = np.zeros(21, dtype=np.float64)
        self._f("hmm_params")(gc, go_h, go_u, pct_id, out.ctypes.data)
        return out

    def hmm_run(self, sym: bytes, params):
        params = np.ascontiguousarray(params, dtype=np.float64)
        pred = np.zeros(max(len(sym), 1), dtype=np.uint8)
        post = np.zeros(max(len(sym), 1), dtype=np.float64)
        r = self._f("hmm_run")(sym, len(sym), params.ctypes.data, pred.ctypes.data, post.ctypes.data)
        if r != 0:
            raise RuntimeError("hmm_run failed")
        return pred[:len(sym)].tobytes(), post[:len(sym)]


def oracle_checker():
    return Checker(oracle(), "orc_")


def ref_checker():
    return Checker(ref(), "ref_")
