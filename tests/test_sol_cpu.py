"""CPU: seed occurrence list + anchor scores.
  * the C restatement (oracle/mauve_oracle.c orc_sol_build / orc_anchor_scores) against the committed golden vectors minted from
    the reference's own SeedOccurrenceList::construct / GetPairwiseAnchorScore (tests/golden/make_golden_sol.py) and, where
    oracle/_ref is built, against the reference directly on further inputs;
  * the __host__ __device__ value functions of csrc/sol.cu, run on the CPU through tests/_emu.py, against the restatement:
    the arithmetic the kernels execute is checked here, the kernels' indexing on the GPU (tests/test_zzzz_next_rows_gpu.py).
"""
import hashlib
import os

import numpy as np
import pytest

import _emu
import _golden
import _oracle
from _bins import NEXT_BIN, run_next, write_fasta
from mauve_py_b200 import synth


def _small():
    z = _golden.npz("sol_small.npz")
    return z, _golden.cases(z)


def test_oracle_sol_and_anchor_scores_match_the_golden_vectors(orc):
    z, cases = _small()
    assert len(cases) >= 12
    for c in cases:
        s0, s1 = z["seq_%s_0" % c["name"]].tobytes(), z["seq_%s_1" % c["name"]].tobytes()
        k = c["key"]
        assert orc.get_seed(c["w"], c["rank"]) == c["seed"]
        f0, f1 = _oracle.sol_build(s0, c["seed"]), _oracle.sol_build(s1, c["seed"])
        assert np.array_equal(f0.view(np.uint32), z[k + "_f0"].view(np.uint32)), k
        assert np.array_equal(f1.view(np.uint32), z[k + "_f1"].view(np.uint32)), k
        rows, _ = orc.find_mums(s0, s1, c["seed"], 0)
        assert np.array_equal(rows, z[k + "_rows"]), k
        for pen, name in ((False, "_lcb"), (True, "_lcb_pen")):
            lcb, ms = _oracle.anchor_scores(s0, s1, c["seed"], rows, z[k + "_cuts"], pen, freq=(f0, f1))
            assert np.array_equal(lcb, z[k + name]), (k, pen)
            assert ms.sum() == lcb.sum()


def test_oracle_sol_mds42_golden(orc):
    """BASELINE config 1: frequencies of both MDS42 genomes (coding seed w15) and the anchor scores of the 29,403-row list"""
    z = _golden.npz("sol_mds42.npz")
    md = _golden.meta(z)
    g0, g1 = _golden.mds42()
    f0, f1 = _oracle.sol_build(g0, md["seed"]), _oracle.sol_build(g1, md["seed"])
    assert hashlib.sha1(f0.tobytes()).hexdigest() == md["sha1_f0"]
    assert hashlib.sha1(f1.tobytes()).hexdigest() == md["sha1_f1"]
    assert np.array_equal(f0[z["idx0"]], z["val0"]) and np.array_equal(f1[z["idx1"]], z["val1"])
    rows = _golden.npz("mums_mds42.npz")["rows_w15_r3"]
    lcb, _ = _oracle.anchor_scores(g0, g1, md["seed"], rows, z["cuts"], False, freq=(f0, f1))
    assert np.array_equal(lcb, z["lcb"])


@pytest.mark.skipif(not _oracle.have_ref_full(), reason="oracle/_ref/libmauve_ref_full.so not built (needs /root/reference)")
@pytest.mark.parametrize("w,rank", [(7, 0), (13, 2), (15, 3), (21, 0), (25, 0)])
def test_oracle_sol_vs_reference_more_inputs(orc, w, rank):
    rng = np.random.default_rng(w)
    a, b = synth.repeat_rich_pair(n=60_000, unit=61, copies=300, seed=w)
    seqs = [a, b, b"A" * 500, bytes(rng.choice(list(b"ACGT"), 31).astype(np.uint8)), bytes(rng.choice(list(b"ACGTN-"), 3000).astype(np.uint8)).replace(b"-", b"n")]
    seed = orc.get_seed(w, rank)
    for s in seqs:
        if len(s) < orc.seed_length(seed):
            continue
        x, y = _oracle.sol_build(s, seed), _oracle.sol_build(s, seed, use_ref=True)
        assert np.array_equal(x.view(np.uint32), y.view(np.uint32)), (len(s), w)
    rows, _ = orc.find_mums(a, b, seed, 0)
    if rows.shape[0]:
        cuts = np.array([0, rows.shape[0] // 3, rows.shape[0]], dtype=np.uint64)
        for pen in (False, True):
            assert np.array_equal(_oracle.anchor_scores(a, b, seed, rows, cuts, pen)[0], _oracle.anchor_scores(a, b, seed, rows, cuts, pen, use_ref=True)[0])


def test_property_checker_sol_expected(orc):
    """the numpy evaluation the full-size GPU test relies on equals the restatement"""
    import _properties as P
    a, _b = synth.repeat_rich_pair(n=50_000, unit=61, copies=400, seed=9)
    for w, rank in ((11, 0), (15, 3), (19, 3)):
        seed = orc.get_seed(w, rank)
        L, wt = orc.seed_length(seed), orc.seed_weight(seed)
        pos, mer = orc.sml_build(a, seed)
        mask = ((1 << 64) - 1) ^ ((1 << (64 - 2 * wt)) - 1)
        want = P.sol_expected(pos, mer, mask, len(a), L)
        assert np.array_equal(want.view(np.uint32), _oracle.sol_build(a, seed).view(np.uint32))


# ---- the CUDA source's value functions on the CPU --------------------------------------------------------------------
def _emu_sol(orc, seq, seed):
    """sorted list from the oracle in the device's key layout (canon << 2 | strand, genome bit 0) -> csrc/sol.cu on the host"""
    L, w = orc.seed_length(seed), orc.seed_weight(seed)
    pos, mer = orc.sml_build(seq, seed)
    keys = ((mer >> np.uint64(64 - 2 * w)) << np.uint64(2)) | (mer & np.uint64(1))
    out = np.zeros(max(len(seq), 1), dtype=np.float32)
    kb = 4 if 2 * w + 2 <= 32 else 8
    _emu.emu().emu_sol(np.ascontiguousarray(keys).ctypes.data, np.ascontiguousarray(pos).ctypes.data, pos.size, len(seq), L, kb, out.ctypes.data)
    return out[:len(seq)]


@pytest.mark.parametrize("w,rank", [(7, 0), (11, 0), (15, 3), (19, 3), (21, 0), (31, 0)])
def test_device_value_functions_sol(orc, w, rank):
    rng = np.random.default_rng(100 + w)
    a, _b = synth.repeat_rich_pair(n=50_000, unit=61, copies=400, seed=w)
    seed = orc.get_seed(w, rank)
    L = orc.seed_length(seed)
    seqs = [a, b"A" * 3000 + a[:2000] + b"AC" * 700, bytes(rng.choice(list(b"ACGT"), L).astype(np.uint8)), bytes(rng.choice(list(b"ACGT"), L + 1).astype(np.uint8)),
            bytes(rng.choice(list(b"ACGT"), L - 1).astype(np.uint8)), b"G"]
    for s in seqs:
        x = _emu_sol(orc, s, seed)
        y = _oracle.sol_build(s, seed)
        assert np.array_equal(x.view(np.uint32), y.view(np.uint32)), (len(s), w, np.flatnonzero(x != y)[:5])


@pytest.mark.parametrize("key_bytes", [4, 8])
def test_device_value_functions_run_lengths_random_sorted_lists(key_bytes):
    """sol_run_length on arbitrary sorted key arrays: runs of every length around the walk limit (8), runs touching both ends of the
    list, one run spanning the whole list; strand bits vary inside a run as they do in a real list"""
    import _properties as P
    rng = np.random.default_rng(17 + key_bytes)
    L = 13
    for trial in range(40):
        npos = int(rng.integers(1, 3000))
        nkeys = int(rng.integers(1, max(2, npos // int(rng.integers(1, 40)))))
        top = (1 << 30) if key_bytes == 4 else (1 << 50)
        canon = np.sort(rng.choice(rng.integers(0, top, size=nkeys), size=npos))
        if trial == 0:
            canon[:] = canon[0]
        keys = (canon.astype(np.uint64) << np.uint64(2)) | rng.integers(0, 2, size=npos).astype(np.uint64)
        keys = np.sort(keys)
        vals = rng.permutation(npos).astype(np.uint32)
        n = npos + L - 1
        out = np.zeros(n, dtype=np.float32)
        _emu.emu().emu_sol(keys.ctypes.data, vals.ctypes.data, npos, n, L, key_bytes, out.ctypes.data)
        want = P.sol_expected(vals, keys, (1 << 64) - 4, n, L)
        assert np.array_equal(out.view(np.uint32), want.view(np.uint32)), (trial, npos, nkeys)


def test_device_value_functions_sol_golden(orc):
    z, cases = _small()
    for c in cases:
        for g in (0, 1):
            s = z["seq_%s_%d" % (c["name"], g)].tobytes()
            x = _emu_sol(orc, s, c["seed"])
            assert np.array_equal(x.view(np.uint32), z["%s_f%d" % (c["key"], g)].view(np.uint32)), c["key"]


def test_device_value_functions_anchor_scores(orc):
    z, cases = _small()
    hoxd = np.array([91, -114, -31, -123, -114, 100, -125, -31, -31, -125, 100, -114, -123, -31, -114, 91], dtype=np.int32)
    other = np.array([5, -4, -4, -4, -4, 5, -4, -4, -4, -4, 5, -4, -4, -4, -4, 5], dtype=np.int32)
    for c in cases:
        k = c["key"]
        s0, s1 = z["seq_%s_0" % c["name"]].tobytes(), z["seq_%s_1" % c["name"]].tobytes()
        rows = np.ascontiguousarray(z[k + "_rows"])
        f0, f1 = np.ascontiguousarray(z[k + "_f0"]), np.ascontiguousarray(z[k + "_f1"])
        for pen in (0, 1):
            for mat in (hoxd, other):
                ms = np.zeros(max(rows.shape[0], 1), dtype=np.int64)
                _emu.emu().emu_anchor_scores(s0, s1, f0.ctypes.data, f1.ctypes.data, rows.ctypes.data, rows.shape[0], mat.ctypes.data, pen, ms.ctypes.data)
                lcb, oms = _oracle.anchor_scores(s0, s1, c["seed"], rows, z[k + "_cuts"], bool(pen), freq=(f0, f1), matrix=mat)
                assert np.array_equal(ms[:rows.shape[0]], oms), (k, pen)
                if mat is hoxd:
                    cuts = z[k + "_cuts"].astype(np.int64)
                    sums = np.array([float(ms[cuts[i]:cuts[i + 1]].sum()) for i in range(cuts.size - 1)])
                    assert np.array_equal(sums, z[k + ("_lcb_pen" if pen else "_lcb")]), (k, pen)


# ---- the C++ adapters' host code next to the reference classes (device calls answered by the restatement) --------------------
@pytest.mark.skipif(not os.path.exists(NEXT_BIN), reason="oracle/_ref/dropin_check_next not built (needs /root/reference at build time)")
def test_adapters_host_code_next_to_the_reference_classes(tmp_path):
    """CudaSeedOccurrenceList / CudaPairwiseAnchorScores (mauve_py_b200/adapters) inside the reference's own flow
    (PairwiseMatchFinder -> EliminateOverlaps_v2 -> IdentifyBreakpoints -> ComputeLCBs_v2 -> GetPairwiseAnchorScore): here the
    device entry points are answered by the CPU restatement through an LD_PRELOAD stub, so this checks the adapters' marshalling;
    the same binary runs against the real library in the GPU suite."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present: the GPU suite runs this binary against the real library")
    a, b = synth.small_pair(200000, seed=51, snp=0.02, n_inv=3)
    write_fasta(tmp_path / "a.fa", "a", a)
    write_fasta(tmp_path / "b.fa", "b", b)
    rc, kv, out = run_next(["sol", tmp_path / "a.fa", 15, 3])
    assert rc == 3 and "no CUDA device" in out           # no CPU fallback behind the adapters
    env = dict(os.environ, LD_PRELOAD=_emu.stub_library())
    rc, kv, out = run_next(["sol", tmp_path / "a.fa", 15, 3], env)
    assert rc == 0 and kv["RESULT"] == "identical" and int(kv["not_one"]) > 0, out
    for w, r in ((15, 3), (11, 0)):
        rc, kv, out = run_next(["scores", tmp_path / "a.fa", tmp_path / "b.fa", w, r], env)
        assert rc == 0 and kv["RESULT"] == "identical" and int(kv["lcbs"]) > 1 and int(kv["reverse_rows"]) > 0, out
