"""TEST INFRASTRUCTURE ONLY: runs the __host__ __device__ value functions of the CUDA sources on the CPU.

csrc/sol.cu keeps everything that decides a value in __host__ __device__ functions; compiled with -DMCU_HOST_EMU the file
gains extern "C" drivers that replace the grid by a loop (and loses its kernels and host API).  The result,
tests/_emu/libmcu_emu.so, lets the CPU suite check that code against the oracle without a GPU.  It is never shipped: the
product library (mauve_py_b200/libmauve_cuda.so) is built without the macro and has no host path.
"""
import ctypes as C
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "mauve_py_b200", "csrc")
OUT = os.path.join(ROOT, "tests", "_emu", "libmcu_emu.so")
SOURCES = ["sol.cu", "dpwild.cu", "hmm.cu", "anchorcols.cu"]

_lib = None


def emu():
    global _lib
    if _lib is not None:
        return _lib
    from mauve_py_b200 import _build
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in deps):
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        cmd = [_build.nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17", "-DMCU_HOST_EMU", "-Xcompiler",
               "-fPIC,-ffp-contract=off", "-diag-suppress", "177", "-shared", "-o", OUT] + srcs
        subprocess.check_call(cmd)
    L = C.CDLL(OUT)
    u64, vp = C.c_uint64, C.c_void_p
    L.emu_sol.argtypes = [vp, vp, u64, u64, C.c_int, C.c_int, vp]
    L.emu_sol.restype = None
    L.emu_anchor_scores.argtypes = [vp, vp, vp, vp, vp, u64, vp, C.c_int, vp]
    L.emu_anchor_scores.restype = None
    L.emu_nw_wild.argtypes = [C.c_char_p, C.c_uint, C.c_char_p, C.c_uint, vp, vp]
    L.emu_nw_wild.restype = C.c_longlong
    L.emu_hmm_chain.argtypes = [vp, u64, vp, C.c_int, vp, vp, vp]
    L.emu_hmm_chain.restype = None
    L.emu_hmm_run.argtypes = [vp, u64, vp, vp, vp, vp]
    L.emu_hmm_run.restype = C.c_int
    L.emu_hmm_vchain.argtypes = [vp, u64, vp, C.c_int, vp]
    L.emu_hmm_vchain.restype = None
    L.emu_hmm_fprod.argtypes = [vp, vp, u64, vp]
    L.emu_hmm_fprod.restype = None
    L.emu_anchor_cols.argtypes = [vp, C.c_uint, C.c_uint, C.c_uint, vp, vp, vp, vp, vp]
    L.emu_anchor_cols.restype = C.c_longlong
    L.emu_anchor_counters.argtypes = [vp]
    L.emu_anchor_counters.restype = None
    _lib = L
    return L


STUB = os.path.join(ROOT, "tests", "_stub", "libmcu_stub.so")


def stub_library():
    """LD_PRELOAD stand-in for libmauve_cuda.so answering from the CPU restatement (tests/_stub/mcu_stub.c): lets the CPU suite run
    the C++ adapters' host code inside oracle/_ref/dropin_check_next without a GPU.  Returns the path (built on demand)."""
    src = os.path.join(ROOT, "tests", "_stub", "mcu_stub.c")
    import _oracle
    _oracle.oracle()  # builds oracle/libmauve_oracle.so when stale
    if not os.path.exists(STUB) or os.path.getmtime(STUB) < max(os.path.getmtime(src), os.path.getmtime(_oracle.ORACLE_SO)):
        subprocess.check_call(["/usr/bin/gcc", "-O2", "-fPIC", "-shared", "-o", STUB, src, "-L" + os.path.dirname(_oracle.ORACLE_SO), "-lmauve_oracle",
                               "-Wl,-rpath," + os.path.dirname(_oracle.ORACLE_SO)])
    return STUB


BENCH_STUB = os.path.join(ROOT, "tests", "_stub", "libmcu_bench_stub.so")


def bench_stub_library():
    """Stand-in for EVERY symbol of libmauve_cuda.so answered from the CPU restatement (tests/_stub/mcu_bench_stub.c): lets the CPU
    suite execute bench.py's and libmems.py's host code end to end (tests/test_bench_dryrun.py).  Returns the path."""
    here = os.path.join(ROOT, "tests", "_stub")
    srcs = [os.path.join(here, "mcu_bench_stub.c"), os.path.join(here, "mcu_stub.c")]
    import _oracle
    _oracle.oracle()
    if not os.path.exists(BENCH_STUB) or os.path.getmtime(BENCH_STUB) < max([os.path.getmtime(s) for s in srcs] + [os.path.getmtime(_oracle.ORACLE_SO)]):
        subprocess.check_call(["/usr/bin/gcc", "-O2", "-fPIC", "-shared", "-o", BENCH_STUB, srcs[0], "-L" + os.path.dirname(_oracle.ORACLE_SO),
                               "-lmauve_oracle", "-Wl,-rpath," + os.path.dirname(_oracle.ORACLE_SO)])
    return BENCH_STUB
