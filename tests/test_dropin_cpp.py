"""Drop-in check in the reference's own language: the reference's C++ classes next to the adapters of
mauve_py_b200/adapters (CudaDNAMemorySML, CudaPairwiseMatchFinder, CudaMemHash, CudaGlobalAlignBatch, run_cuda), all in one
process, on the same inputs (oracle/dropin_check.cpp, built by oracle/Makefile.ref into oracle/_ref/dropin_check where
/root/reference exists; the binary travels to the GPU box)."""
import gzip
import os
import subprocess

import numpy as np
import pytest

import _golden
from mauve_py_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "dropin_check")


def _run(*args, timeout=900):
    r = subprocess.run([BIN] + [str(a) for a in args], capture_output=True, text=True, timeout=timeout)
    kv = dict(l.split(" ", 1) for l in r.stdout.splitlines() if " " in l)
    return r.returncode, kv, r.stdout + r.stderr


def _fasta(path, name, seq):
    with open(path, "wb") as f:
        f.write(b">" + name.encode() + b"\n")
        for i in range(0, len(seq), 80):
            f.write(seq[i:i + 80] + b"\n")


needs_bin = pytest.mark.skipif(not os.path.exists(BIN), reason="oracle/_ref/dropin_check not built (needs /root/reference at build time)")


@needs_bin
def test_dropin_refuses_without_device():
    """no CPU fallback behind the adapters: without a CUDA device the binary stops with the library's error"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    rc, kv, out = _run("hmm", 100, 1)
    assert rc == 3 and "no CUDA device" in out


@needs_bin
@pytest.mark.gpu
def test_dropin_mds42_buildindex_pair(tmp_path):
    """BASELINE config 1: the MDS42 pair with the default coding seed, as `progressiveMauve --mums` would list it"""
    g0, g1 = _golden.mds42()
    _fasta(tmp_path / "recoded.fa", "recoded", g0)
    _fasta(tmp_path / "full.fa", "full", g1)
    rc, kv, out = _run("mums", tmp_path / "recoded.fa", tmp_path / "full.fa", 0, 3)
    assert rc == 0 and kv["RESULT"] == "identical", out
    assert kv["matches_reference"] == kv["matches_cuda"] == "29403" and kv["sum_len"] == "3792460" and kv["reverse"] == "1515", out
    assert kv["collisions"].split() == ["2714688", "2714688"], out


@needs_bin
@pytest.mark.gpu
@pytest.mark.parametrize("w,r,mode", [(11, 0, ""), (15, 3, "memhash"), (19, 3, ""), (9, 0, "memhash")])
def test_dropin_mums_synthetic(tmp_path, w, r, mode):
    a, b = synth.small_pair(300000, seed=40 + w, snp=0.02, n_inv=3)
    _fasta(tmp_path / "a.fa", "a", a)
    _fasta(tmp_path / "b.fa", "b", b)
    args = ["mums", tmp_path / "a.fa", tmp_path / "b.fa", w, r] + ([mode] if mode else [])
    rc, kv, out = _run(*args)
    assert rc == 0 and kv["RESULT"] == "identical" and int(kv["matches_cuda"]) > 100, out


@needs_bin
@pytest.mark.gpu
@pytest.mark.parametrize("w,r", [(0, 3), (7, 0), (21, 0), (31, 0)])
def test_dropin_sml(tmp_path, w, r):
    a, _ = synth.small_pair(500000, seed=60 + w)
    _fasta(tmp_path / "a.fa", "a", a)
    rc, kv, out = _run("sml", tmp_path / "a.fa", w, r)
    assert rc == 0 and kv["RESULT"] == "identical", out


@needs_bin
@pytest.mark.gpu
def test_dropin_gap_search_batch():
    """one round of recursive anchoring: 2000 gap pairs through the reference's per-gap loop and through one batched call"""
    rc, kv, out = _run("gaps", 2000, 11)
    assert rc == 0 and kv["RESULT"] == "identical" and int(kv["matches"]) > 2000, out


@needs_bin
@pytest.mark.gpu
def test_dropin_dp_and_hmm():
    rc, kv, out = _run("dp", 120, 7)
    assert rc == 0 and kv["RESULT"] == "identical", out
    rc, kv, out = _run("hmm", 200000, 3)
    assert rc == 0 and kv["RESULT"] == "identical" and int(kv["homologous_columns"]) > 1000, out


@needs_bin
@pytest.mark.gpu
def test_dropin_hmm_lcb_sized_string_is_bit_faithful():
    """3 M columns: an evaluation in double is 6.5e-5 away from the reference's float32 bfloat arithmetic here and flips H/N
    calls; the bfloat-faithful chains reproduce run() exactly"""
    rc, kv, out = _run("hmm", 3000000, 5)
    assert rc == 0 and kv["RESULT"] == "identical" and kv["differing_columns"] == "0" and kv["threshold_columns"] == "0", out
    assert float(kv["max_rel_err"]) < 1e-12, out


@needs_bin
def test_dropin_adapters_host_code_through_the_stub(tmp_path):
    """the adapters' own host code (sequence extraction, Match construction, PWPath conversion) next to the reference classes on a
    machine without a GPU: the device entry points are answered by the CPU restatement through an LD_PRELOAD stub (tests/_stub).
    The GPU suite runs the same binary against the real library."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present: the GPU tests above run this binary against the real library")
    import _emu
    env = dict(os.environ, LD_PRELOAD=_emu.stub_library())

    def run(*args):
        r = subprocess.run([BIN] + [str(a) for a in args], capture_output=True, text=True, timeout=600, env=env)
        return r.returncode, dict(l.split(" ", 1) for l in r.stdout.splitlines() if " " in l), r.stdout + r.stderr

    a, b = synth.small_pair(200000, seed=52, snp=0.02, n_inv=3)
    _fasta(tmp_path / "a.fa", "a", a)
    _fasta(tmp_path / "b.fa", "b", b)
    for w, r in ((15, 3), (21, 0)):
        rc, kv, out = run("sml", tmp_path / "a.fa", w, r)
        assert rc == 0 and kv["RESULT"] == "identical", out
    for args in ((15, 3), (9, 0, "memhash")):
        rc, kv, out = run("mums", tmp_path / "a.fa", tmp_path / "b.fa", *args)
        assert rc == 0 and kv["RESULT"] == "identical" and int(kv["matches_cuda"]) > 100, out
    rc, kv, out = run("dp", 40, 7)
    assert rc == 0 and kv["RESULT"] == "identical", out
