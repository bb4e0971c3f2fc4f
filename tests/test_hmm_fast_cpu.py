"""CPU check of the FP32 form of the HomologyHMM recurrence (csrc/hmm.cu: hmm_fprod / hmm_fast_step / the hazard test).

The kernels replace the reference's y = (float)((double)v * c) (bfloat_pr_double_product, algebras.h:225-231) by one FMUL and one
FFMA wherever a hazard test says the result is provably the same float, and fall back to the operation-by-operation form elsewhere.
tests/_emu.py compiles exactly those __host__ __device__ functions for the CPU (never shipped); here they face
  * the definition itself on random and on ADVERSARIAL operands (products placed within a few double ulps of a float rounding
    boundary, where double rounding and single rounding part ways), and
  * the oracle (orc_hmm_run = the reference's run(), pinned on oracle/_ref) on whole strings, posterior for posterior.
"""
import ctypes as C

import numpy as np
import pytest

import _emu
import _oracle


def _counts():
    return np.zeros(8, dtype=np.uint64)


def _fprod(v, c):
    v = np.ascontiguousarray(v, dtype=np.float32)
    c = np.ascontiguousarray(c, dtype=np.float64)
    k = _counts()
    _emu.emu().emu_hmm_fprod(v.ctypes.data, c.ctypes.data, int(v.size), k.ctypes.data)
    return k  # accepted, hazardous, accepted-and-different, hazardous-and-different


def _params(**kw):
    return _oracle.oracle_checker().hmm_params(**kw)


def test_fprod_random_operands():
    rng = np.random.default_rng(7)
    n = 20_000_000
    v = np.exp(rng.uniform(np.log(1e-17), np.log(1e17), n)).astype(np.float32)
    # coefficients of the size the model has (transition * emission: 1e-9 .. 1) with full 53-bit mantissas
    c = np.exp(rng.uniform(np.log(1e-9), 0.0, n))
    k = _fprod(v, c)
    assert int(k[2]) == 0, "an accepted FP32 product differs from (float)((double)v * c)"
    assert int(k[0]) + int(k[1]) == n
    # the hazard test must stay cheap: about 2^-14 of the products
    assert int(k[1]) < n // 4000


def test_fprod_at_rounding_boundaries():
    """products steered onto float midpoints: v * c within a few ulps OF THE DOUBLE of (k + 1/2) ulp32 -- exactly where
    (float)(double) and a correctly rounded product disagree.  Every such pair must be flagged or equal."""
    rng = np.random.default_rng(11)
    n = 2_000_000
    v = np.exp(rng.uniform(np.log(1e-6), np.log(1e6), n)).astype(np.float32)
    y = np.exp(rng.uniform(np.log(1e-10), np.log(1e10), n)).astype(np.float32)
    mid = (y.astype(np.float64) + np.nextafter(y, np.float32(np.inf)).astype(np.float64)) / 2.0   # exact in double
    c = mid / v.astype(np.float64)
    steps = rng.integers(-3, 4, n)
    for _ in range(3):
        c = np.where(steps > 0, np.nextafter(c, np.inf), np.where(steps < 0, np.nextafter(c, -np.inf), c))
        steps = steps - np.sign(steps)
    k = _fprod(v, c)
    assert int(k[2]) == 0
    assert int(k[1]) > n * 0.99          # they ARE hazards
    assert int(k[3]) > 0                 # and some of them would indeed have come out differently: the test has teeth
    # powers of two (the boundary below r sits at a quarter ulp) and exact products
    v2 = np.exp2(rng.integers(-40, 40, 100000)).astype(np.float32)
    c2 = np.exp2(rng.integers(-20, 0, 100000).astype(np.float64)) * (1 + np.exp2(-rng.integers(24, 53, 100000).astype(np.float64)))
    assert int(_fprod(v2, c2)[2]) == 0
    c3 = np.exp2(rng.integers(-20, 0, 100000).astype(np.float64)) * (1 - np.exp2(-rng.integers(24, 54, 100000).astype(np.float64)))
    assert int(_fprod(v2, c3)[2]) == 0


def _random_string(rng, n, p_match=0.7, blocks=True):
    """column symbols '1'..'8' (encoder LM/Islands.h:90-155): stretches that look homologous and stretches that do not"""
    if not blocks:
        return (rng.integers(0, 8, n) + ord("1")).astype(np.uint8)
    out = np.empty(n, dtype=np.uint8)
    pos = 0
    while pos < n:
        ln = int(rng.integers(50, 5000))
        hom = rng.random() < p_match
        pr = np.array([.35, .3, .03, .06, .03, .03, .1, .1]) if hom else np.array([.12, .12, .12, .12, .12, .12, .14, .14])
        pr = pr / pr.sum()
        out[pos:pos + ln] = (rng.choice(8, size=min(ln, n - pos), p=pr) + ord("1")).astype(np.uint8)
        pos += ln
    return out


@pytest.mark.parametrize("fwd", [1, 0])
def test_float_step_equals_exact_step_along_a_chain(fwd):
    """hmm_float_step (what the re-examination and the thread-per-string kernel evaluate a column with) against the operation-by-
    operation step at every column of a 3 M-column chain: identical wherever it does not report a hazard"""
    rng = np.random.default_rng(3 + fwd)
    p = _params()
    n = 3_000_000
    sym = _random_string(rng, n)
    f = np.zeros(n, dtype=np.float32)
    e = np.zeros(n, dtype=np.int32)
    k = _counts()
    _emu.emu().emu_hmm_chain(sym.ctypes.data, n, p.ctypes.data, fwd, f.ctypes.data, e.ctypes.data, k.ctypes.data)
    assert int(k[2]) == 0, "an accepted FP32 step differs from the operation-by-operation step"
    assert int(k[1]) + int(k[3]) == n - 1
    assert 0 < int(k[3]) < n // 300          # hazards exist and are rare
    assert e.min() < -1000                   # the string renormalises thousands of times


@pytest.mark.parametrize("kw", [dict(), dict(gc=0.35), dict(gc=0.65, go_h=0.0005, go_u=0.00002), dict(pct_id=0.8)])
def test_virtual_chain_carries_the_reference_values(kw):
    """the chain of the long-string kernel: both states on one running scale, no exponents (hmm_vstep + the rescaling rule), beside
    the exact chain.  Its VALUES are the reference's at every column but the rare rounding hazards (where the kernel repairs the
    chain); the reference's own split into mantissa and exponent is what hmm_canon gives, or another form of the same value"""
    p = _params(**kw)
    for seed, block in ((1, 3000), (2, 300), (3, 40)):
        from mauve_py_b200 import synth
        sym = np.frombuffer(synth.hmm_string(1_000_000, seed=seed, block=block), dtype=np.uint8).copy()
        for fwd in (1, 0):
            k = _counts()
            _emu.emu().emu_hmm_vchain(sym.ctypes.data, sym.size, p.ctypes.data, fwd, k.ctypes.data)
            same_form, other_form, value_diff, rescales, tiny = (int(x) for x in k[:5])
            assert same_form + other_form + value_diff == sym.size - 1
            assert value_diff <= 2, (kw, seed, fwd, value_diff)          # rounding hazards: about one column in 20 million
            assert other_form < sym.size // 100                          # handed on by the re-examination: one pass more for that block
            assert rescales > sym.size // 60 and tiny == 0               # the scale moves every ~39 columns; nothing near the denormals


@pytest.mark.parametrize("kw", [dict(), dict(gc=0.35), dict(gc=0.65, go_h=0.0005, go_u=0.00002), dict(pct_id=0.8)])
def test_run_equals_oracle(kw):
    """the kernels' whole value path on the CPU against the reference's run(): identical posteriors (as doubles) and calls"""
    rng = np.random.default_rng(101)
    p = _params(**kw)
    orc = _oracle.oracle_checker()
    for n, blocks in ((1, True), (2, True), (33, False), (1000, False), (300_000, True)):
        sym = _random_string(rng, n, blocks=blocks)
        pred = np.zeros(n, dtype=np.uint8)
        post = np.zeros(n, dtype=np.float64)
        k = _counts()
        assert _emu.emu().emu_hmm_run(sym.ctypes.data, n, p.ctypes.data, pred.ctypes.data, post.ctypes.data, k.ctypes.data) == 0
        want_pred, want_post = orc.hmm_run(sym.tobytes(), p)
        assert int(k[2]) == 0
        assert bytes(pred) == bytes(want_pred)
        assert np.array_equal(post.view(np.uint64), np.asarray(want_post, dtype=np.float64).view(np.uint64))


def test_run_equals_golden_mds42_call():
    """the one call progressiveMauve really makes on the MDS42 pair (first million columns, tests/golden/hmm_mds42_call.npz)"""
    import os
    path = os.path.join(os.path.dirname(__file__), "golden", "hmm_mds42_call.npz")
    z = np.load(path)
    keys = set(z.files)
    sym = np.frombuffer(z["sym"].tobytes(), dtype=np.uint8).copy()
    n = int(sym.size)
    p = np.ascontiguousarray(z["params"], dtype=np.float64) if "params" in keys else _params()
    pred = np.zeros(n, dtype=np.uint8)
    k = _counts()
    assert _emu.emu().emu_hmm_run(sym.ctypes.data, n, p.ctypes.data, pred.ctypes.data, None, k.ctypes.data) == 0
    assert int(k[2]) == 0
    want = np.frombuffer(z["pred"].tobytes(), dtype=np.uint8)
    assert np.array_equal(pred, want)
