"""CPU: the C-ABI library loads and exports every symbol include/mauve_cuda.h declares; host-only entry points work
without a GPU; compute entry points fail loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import _golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu():
    return os.path.exists("/dev/nvidiactl") or os.path.exists("/dev/nvidia0")


def test_library_exports_header_symbols():
    import mauve_py_b200 as mp
    hdr = open(os.path.join(ROOT, "include", "mauve_cuda.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(mcu_[a-z0-9_]+)\s*\(", hdr)))
    assert declared == sorted(mp.SYMBOLS)
    lib = mp.lib()
    for name in declared:
        assert hasattr(lib, name), name


def test_seed_tables_host_side():
    import mauve_py_b200 as mp
    t = _golden.seeds()
    for k, v in t["get_seed"].items():
        w, r = (int(x) for x in k.split(","))
        assert mp.getSeed(w, r) == v, k
    for k, v in t["seed_length"].items():
        assert mp.getSeedLength(int(k)) == v
    for k, v in t["seed_weight"].items():
        assert mp.getSeedWeight(int(k)) == v
    for k, v in t["default_weight"].items():
        assert mp.getDefaultSeedWeight(int(k)) == v


def test_hmm_params_host_side():
    import mauve_py_b200 as mp
    z = _golden.npz("hmm_small.npz")
    for c in _golden.cases(z):
        assert np.array_equal(mp.libmems.hmm_params(c[1], c[2], c[3], c[4]), z["params%d" % c[0]])
    p = mp.getAdaptedHoxdMatrixParameters(0.5)
    q = mp.adaptToPercentIdentity(p, 0.7)
    assert np.allclose(q.as_array(), mp.libmems.hmm_params(0.5, 0, 0, 0.7), rtol=0, atol=0)


@pytest.mark.skipif(_has_gpu(), reason="a GPU is present: the no-device behaviour cannot be observed")
def test_compute_calls_fail_loudly_without_gpu():
    import mauve_py_b200 as mp
    from mauve_py_b200 import _capi
    with pytest.raises(mp.McuError) as e:
        mp.DNAMemorySML().Create(b"ACGTACGTACGTACGTACGT", mp.getSeed(5, 0))
    assert e.value.code == _capi.MCU_ENODEV
    with pytest.raises(mp.McuError):
        mp.GlobalAlign(b"ACGT", b"ACGT")
    with pytest.raises(mp.McuError):
        mp.run(b"1234", mp.libmems.hmm_params())
    with pytest.raises(mp.McuError):
        mp.libmems.find_mums(b"ACGT" * 10, b"ACGT" * 10, mp.getSeed(5, 0))
    # the entry points added for the rows next to the path and for wildcard regions: no host path behind them either
    sml = mp.DNAMemorySML()
    sml._seq, sml._seed, sml._length = np.frombuffer(b"ACGT" * 10, dtype=np.uint8), mp.getSeed(5, 0), 40
    with pytest.raises(mp.McuError) as e:
        mp.SeedOccurrenceList().construct(sml)
    assert e.value.code == _capi.MCU_ENODEV
    with pytest.raises(mp.McuError) as e:
        mp.libmems.anchor_scores(b"ACGT" * 10, b"ACGT" * 10, np.array([[8, 1, 1]], dtype=np.int64), [0, 1], seed=mp.getSeed(5, 0))
    assert e.value.code == _capi.MCU_ENODEV
    with pytest.raises(mp.McuError) as e:
        mp.GlobalAlignBatchWild([(b"ACGN", b"ACGT")])
    assert e.value.code == _capi.MCU_ENODEV


def test_product_does_not_touch_oracle():
    """the package never imports, links or opens anything under oracle/"""
    pkg = os.path.join(ROOT, "mauve_py_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "libmauve_oracle" not in txt and "libmauve_ref" not in txt and "mauve_oracle.c" not in txt, f
                assert not re.search(r"^\s*(from|import)\s+_oracle", txt, flags=re.M), f


def test_ctypes_prototypes_match_the_header_arity():
    """every prototype of include/mauve_cuda.h has as many parameters as the ctypes binding declares (no compute call)"""
    import re
    import mauve_py_b200 as mp
    lib = mp.lib()
    text = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "mauve_cuda.h")).read(), flags=re.S)
    protos = re.findall(r"\b(mcu_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text)
    assert len(protos) == len(mp.SYMBOLS) and {n for n, _ in protos} == set(mp.SYMBOLS)
    for name, args in protos:
        args = args.strip()
        n = 0 if args in ("", "void") else len(args.split(","))
        declared = getattr(lib, name).argtypes
        assert (declared is None and n == 0) or (declared is not None and len(declared) == n), (name, n, declared)
