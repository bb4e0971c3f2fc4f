#!/usr/bin/env python
"""A/B aid: builds build/variants/libmauve_cuda_<name>.so = this tree with one patch of experiments/ applied (the tree itself is not
touched).  The variant travels to the GPU box with the snapshot (build/ is git-ignored, not gpurun-ignored) and is selected with
MAUVE_CUDA_LIB=<path> (mauve_py_b200/_capi.py).      python tools/build_variant.py experiments/<x>.patch <name>"""
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
patch, name = sys.argv[1], sys.argv[2]
work = tempfile.mkdtemp()
try:
    for d in ("mauve_py_b200", "include"):
        shutil.copytree(os.path.join(ROOT, d), os.path.join(work, d), ignore=shutil.ignore_patterns("*.so", "__pycache__"))
    subprocess.check_call(["patch", "-p0", "-i", os.path.abspath(patch)], cwd=work)
    sys.path.insert(0, work)
    env = dict(os.environ, PYTHONPATH=work)
    subprocess.check_call([sys.executable, "-c", "from mauve_py_b200 import _build; print(_build.build_library(force=True))"], cwd=work, env=env)
    out = os.path.join(ROOT, "build", "variants")
    os.makedirs(out, exist_ok=True)
    dst = os.path.join(out, "libmauve_cuda_%s.so" % name)
    shutil.copy(os.path.join(work, "mauve_py_b200", "libmauve_cuda.so"), dst)
    print(dst)
finally:
    shutil.rmtree(work, ignore_errors=True)
