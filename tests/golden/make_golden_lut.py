"""Mints tests/golden/mds42_lut.npz: the int32 index LUT of `mauve.buildIndex(mds42_recoded.fa, mds42_full.fa)`, computed by the
REFERENCE ITSELF in the build container (it cannot travel, the vectors do):

  * the Python layer is the reference's own, imported from /root/reference (mauve/__init__.py, buildindex.py): `buildIndex`
    runs `progressiveMauveStatic`, parses the XMFA and builds the LUT (buildindex.py:90-138);
  * `progressiveMauveStatic` is oracle/_ref/progressiveMauve, the reference's C++ sources compiled in place (oracle/Makefile.ref);
  * `mauve.indexutils` is the reference's Cython file compiled into oracle/_ref/pyref.  Today's Cython rejects one type name in a
    function buildIndex never calls (`np.int_t`, indexutils.pyx:183), so the build reads the .pyx through a one-word sed into that
    scratch directory; nothing of it is kept in the repository;
  * `libnano` (un-vendored, not installable offline: README.md:63-65) only supplies FASTA/XMFA parsing; it is replaced by the
    reference's own in-tree equivalents mauve/fasta.py and mauve/xmfa.py, and `bitarray` (imported, unused by
    buildIndex) by an empty stand-in.

Also records the sha1 of the XMFA body and of the match list, so that the end-to-end test can tell where a difference starts.

    python tests/golden/make_golden_lut.py
"""
import gzip
import hashlib
import importlib.machinery
import importlib.util
import os
import shutil
import subprocess
import sys
import sysconfig
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
PYREF = os.path.join(ROOT, "oracle", "_ref", "pyref")
BINARY = os.path.join(ROOT, "oracle", "_ref", "progressiveMauve")


def build_indexutils():
    os.makedirs(PYREF, exist_ok=True)
    pyx = os.path.join(PYREF, "indexutils.pyx")
    with open(os.path.join(REF, "mauve", "indexutils.pyx")) as f:
        src = f.read().replace("np.int_t", "np.int64_t")
    with open(pyx, "w") as f:
        f.write(src)
    subprocess.check_call([sys.executable, "-m", "cython", "-3", pyx, "-o", os.path.join(PYREF, "indexutils.c")])
    so = os.path.join(PYREF, "indexutils" + sysconfig.get_config_var("EXT_SUFFIX"))
    subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-w", "-I" + sysconfig.get_paths()["include"], "-I" + np.get_include(),
                           os.path.join(PYREF, "indexutils.c"), "-o", so])
    return so


def import_reference_mauve(so, mauve_dir):
    os.environ["MAUVE_DIR"] = mauve_dir
    sys.modules["bitarray"] = types.ModuleType("bitarray")
    sys.path.insert(0, REF)
    # libnano.fileio stand-ins made of the reference's own parsers
    def load(name, path):
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod
    xmfa = load("_ref_xmfa", os.path.join(REF, "mauve", "xmfa.py"))
    fasta = load("_ref_fasta", os.path.join(REF, "mauve", "fasta.py"))
    libnano = types.ModuleType("libnano")
    fileio = types.ModuleType("libnano.fileio")
    m_fasta = types.ModuleType("libnano.fileio.fasta")
    m_xmfa = types.ModuleType("libnano.fileio.xmfa")
    m_fasta.parseFasta = fasta.parseFasta
    m_xmfa.parseXMFA = xmfa.parseXMFA

    def getSeqFromFile(fp):  # libnano's convenience: the sequence of the first record
        return fasta.parseFasta(fp)[0][1]
    fileio.getSeqFromFile = getSeqFromFile
    fileio.fasta, fileio.xmfa, libnano.fileio = m_fasta, m_xmfa, fileio
    sys.modules.update({"libnano": libnano, "libnano.fileio": fileio, "libnano.fileio.fasta": m_fasta, "libnano.fileio.xmfa": m_xmfa})
    loader = importlib.machinery.ExtensionFileLoader("mauve.indexutils", so)
    spec = importlib.util.spec_from_loader("mauve.indexutils", loader)
    iu = importlib.util.module_from_spec(spec)
    loader.exec_module(iu)
    sys.modules["mauve.indexutils"] = iu
    import mauve  # the reference package
    mauve.indexutils = iu
    return mauve


def main():
    assert os.path.exists(BINARY), "run `make -f oracle/Makefile.ref` first"
    so = build_indexutils()
    work = tempfile.mkdtemp()
    try:
        os.symlink(BINARY, os.path.join(work, "progressiveMauveStatic"))
        mauve = import_reference_mauve(so, work)
        fas = []
        for name in ("mds42_recoded", "mds42_full"):
            p = os.path.join(work, name + ".fa")
            with gzip.open(os.path.join(HERE, name + ".fa.gz"), "rb") as f, open(p, "wb") as g:
                shutil.copyfileobj(f, g)
            fas.append(p)
        lut = mauve.buildIndex(fas[0], fas[1])
        assert lut.dtype == np.int32
        # the intermediate artefacts of the same run, for diagnosis: XMFA body and --mums list
        run = os.path.join(work, "run")
        os.makedirs(run)
        subprocess.check_call([BINARY, "--output=mds42.xmfa", "../mds42_recoded.fa", "../mds42_full.fa"], cwd=run, stdout=subprocess.DEVNULL,
                              stderr=subprocess.DEVNULL)
        # header lines name files and the '>' lines end with the FASTA path: hash coordinates, strands and sequences only
        body = b"".join(b" ".join(l.split()[:3]) + b"\n" if l.startswith(b">") else l
                        for l in open(os.path.join(run, "mds42.xmfa"), "rb") if not l.startswith(b"#"))
        subprocess.check_call([BINARY, "--mums", "--output=mds42.mums", "../mds42_recoded.fa", "../mds42_full.fa"], cwd=run,
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        rows = [l.split(b"\t")[:3] for l in open(os.path.join(run, "mds42.mums"), "rb").read().split(b"\n")[7:] if l.strip()]
        meta = {"xmfa_body_sha1": hashlib.sha1(body).hexdigest(), "xmfa_body_bytes": len(body),
                "mums_rows": len(rows), "mums_sha1": hashlib.sha1(b"\n".join(b"\t".join(r) for r in rows)).hexdigest(),
                "lut_sha1": hashlib.sha1(lut.tobytes()).hexdigest(), "mapped": int((lut >= 0).sum()), "length": int(lut.size)}
        # the LUT is piecewise linear: store first differences (compresses to a few hundred kB)
        d = np.diff(lut.astype(np.int64), prepend=0).astype(np.int32)
        np.savez_compressed(os.path.join(HERE, "mds42_lut.npz"), lut_diff=d, meta=np.array(repr(meta)))
        print(meta)
    finally:
        shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
