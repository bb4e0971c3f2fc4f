"""ctypes binding of the C ABI declared in include/mauve_cuda.h (libmauve_cuda.so).

There is no CPU fallback anywhere in this package: if the shared library is missing the import
of `lib()` raises, and if no sm_100 device is usable every compute call raises McuError
(MCU_ENODEV).  Nothing here imports or executes anything under oracle/.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MAUVE_CUDA_LIB") or os.path.join(HERE, "libmauve_cuda.so")   # the override selects an A/B build of the SAME library

MCU_OK, MCU_ENODEV, MCU_ECUDA, MCU_EINVAL, MCU_EGAP, MCU_ENOMEM, MCU_EALPHA, MCU_ESMALL = 0, -1, -2, -3, -4, -5, -6, -7
COMM_ID_BYTES = 128
SOLID_SEED = 0x7FFFFFFF
CODING_SEED = 3
RULE_PAIRWISE, RULE_MEMHASH = 0, 1

# every symbol include/mauve_cuda.h declares (tests check the library exports exactly these)
SYMBOLS = [
    "mcu_init", "mcu_shutdown", "mcu_last_error", "mcu_free", "mcu_host_alloc", "mcu_host_free",
    "mcu_get_seed", "mcu_default_seed_weight", "mcu_seed_length", "mcu_seed_weight",
    "mcu_sml_build", "mcu_sml_last_stats", "mcu_find_mums", "mcu_find_mums_into", "mcu_find_mums_batch",
    "mcu_session_create", "mcu_session_destroy", "mcu_session_upload", "mcu_session_upload_begin", "mcu_session_run",
    "mcu_session_enumerate", "mcu_session_uniq_bitmap", "mcu_session_finish", "mcu_session_merge",
    "mcu_session_match_count", "mcu_session_download", "mcu_session_matches_device",
    "mcu_session_launch_count", "mcu_merge_matches",
    "mcu_nw_batch", "mcu_nw_batch_wild", "mcu_nw_last_stats", "mcu_hmm_params", "mcu_hmm_batch", "mcu_sol_build", "mcu_anchor_scores",
    "mcu_comm_unique_id", "mcu_comm_init", "mcu_comm_destroy", "mcu_comm_rank", "mcu_comm_world", "mcu_comm_barrier",
    "mcu_comm_allreduce_f64", "mcu_comm_gather_bytes", "mcu_device_synchronize",
    "mcu_session_upload_sharded", "mcu_session_run_sharded", "mcu_find_mums_sharded",
    "mcu_test_sort_pairs", "mcu_test_int32_peak", "mcu_test_hmm_counters", "mcu_eliminate_overlaps", "mcu_lcbs", "mcu_sml_build_shard", "mcu_sml_build_sharded",
    "mcu_anchor_default_params", "mcu_anchor_cols_batch", "mcu_test_anchor_counters",
]


class McuError(RuntimeError):
    def __init__(self, code, text):
        super().__init__("libmauve_cuda error %d: %s" % (code, text))
        self.code = code


class Match(C.Structure):
    _fields_ = [("len", C.c_int64), ("start0", C.c_int64), ("start1", C.c_int64)]


_lib = None


def lib():
    """Loads libmauve_cuda.so (building is __graft_entry__.build()'s job); raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    u64, vp, i32 = C.c_uint64, C.c_void_p, C.c_int
    L.mcu_init.argtypes = [i32]
    L.mcu_shutdown.restype = None
    L.mcu_last_error.restype = C.c_char_p
    L.mcu_free.argtypes = [vp]
    L.mcu_free.restype = None
    L.mcu_host_alloc.argtypes = [C.POINTER(vp), u64]
    L.mcu_host_free.argtypes = [vp]
    L.mcu_host_free.restype = None
    L.mcu_get_seed.argtypes = [i32, i32]
    L.mcu_get_seed.restype = u64
    L.mcu_default_seed_weight.argtypes = [u64]
    L.mcu_default_seed_weight.restype = C.c_uint
    L.mcu_seed_length.argtypes = [u64]
    L.mcu_seed_weight.argtypes = [u64]
    L.mcu_sml_build.argtypes = [vp, u64, u64, vp, vp, vp, C.POINTER(u64)]
    L.mcu_sml_last_stats.argtypes = [vp]
    L.mcu_sml_last_stats.restype = None
    L.mcu_find_mums.argtypes = [vp, u64, vp, u64, u64, i32, C.POINTER(C.POINTER(Match)), C.POINTER(u64), vp]
    L.mcu_find_mums_into.argtypes = [vp, u64, vp, u64, u64, i32, vp, u64, C.POINTER(u64), vp]
    L.mcu_find_mums_sharded.argtypes = [vp, u64, vp, u64, u64, i32, vp, u64, C.POINTER(u64), vp]
    L.mcu_session_upload_begin.argtypes = [vp, vp, u64, vp, u64, i32]
    L.mcu_session_upload_sharded.argtypes = [vp, vp, u64, vp, u64]
    L.mcu_session_run_sharded.argtypes = [vp, u64, vp, vp]
    L.mcu_comm_unique_id.argtypes = [vp]
    L.mcu_comm_init.argtypes = [i32, i32, vp]
    L.mcu_comm_destroy.restype = None
    L.mcu_comm_allreduce_f64.argtypes = [vp, i32, i32]
    L.mcu_comm_gather_bytes.argtypes = [vp, u64, C.POINTER(vp), vp]
    L.mcu_find_mums_batch.argtypes = [u64, vp, vp, vp, vp, vp, i32, C.POINTER(C.POINTER(Match)), vp, vp]
    L.mcu_session_create.argtypes = [C.POINTER(vp)]
    L.mcu_session_destroy.argtypes = [vp]
    L.mcu_session_destroy.restype = None
    L.mcu_session_upload.argtypes = [vp, vp, u64, vp, u64]
    L.mcu_session_run.argtypes = [vp, u64, i32, i32, vp, vp]
    L.mcu_session_enumerate.argtypes = [vp, u64, i32, i32]
    L.mcu_session_uniq_bitmap.argtypes = [vp, C.POINTER(vp), C.POINTER(u64)]
    L.mcu_session_finish.argtypes = [vp, i32, vp, vp]
    L.mcu_session_merge.argtypes = [vp, vp, u64, i32, vp]
    L.mcu_session_match_count.argtypes = [vp]
    L.mcu_session_match_count.restype = u64
    L.mcu_session_download.argtypes = [vp, vp]
    L.mcu_session_matches_device.argtypes = [vp]
    L.mcu_session_matches_device.restype = vp
    L.mcu_session_launch_count.argtypes = [vp]
    L.mcu_session_launch_count.restype = u64
    L.mcu_merge_matches.argtypes = [vp, u64, i32, C.POINTER(C.POINTER(Match)), C.POINTER(u64), C.POINTER(u64)]
    L.mcu_nw_batch.argtypes = [u64, vp, vp, vp, vp, vp, vp, vp, vp, C.POINTER(C.c_float)]
    L.mcu_nw_batch_wild.argtypes = [u64, vp, vp, vp, vp, vp, vp, vp, vp, C.POINTER(C.c_float)]
    L.mcu_nw_last_stats.argtypes = [vp]
    L.mcu_nw_last_stats.restype = None
    L.mcu_hmm_params.argtypes = [C.c_double, C.c_double, C.c_double, C.c_double, vp]
    L.mcu_hmm_batch.argtypes = [u64, vp, vp, vp, vp, vp, C.POINTER(C.c_float)]
    L.mcu_sol_build.argtypes = [vp, u64, u64, vp]
    L.mcu_anchor_scores.argtypes = [vp, u64, vp, u64, u64, vp, vp, vp, u64, vp, u64, vp, i32, vp, vp]
    L.mcu_test_sort_pairs.argtypes = [vp, vp, u64, i32, i32]
    L.mcu_test_int32_peak.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_float)]
    L.mcu_test_hmm_counters.argtypes = [C.c_void_p]
    L.mcu_eliminate_overlaps.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
    L.mcu_lcbs.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.mcu_sml_build_shard.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.mcu_sml_build_sharded.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
    L.mcu_anchor_default_params.argtypes = [C.c_void_p]
    L.mcu_anchor_default_params.restype = None
    L.mcu_anchor_cols_batch.argtypes = [C.c_uint64] + [C.c_void_p] * 13
    L.mcu_test_anchor_counters.argtypes = [C.c_void_p]
    _lib = L
    return L


def check(code):
    if code != MCU_OK:
        raise McuError(code, (lib().mcu_last_error() or b"").decode("utf-8", "replace"))
