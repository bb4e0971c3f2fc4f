#!/usr/bin/env python
"""DEVELOPMENT TOOL (not part of the test suite, not used by the product): runs pytest with libmauve_cuda.so replaced by the
oracle-backed stand-in of tests/_stub/mcu_bench_stub.c, so that the PYTHON and host-side code of the `-m gpu` tests (fixtures,
marshalling, child binaries through LD_PRELOAD, assertions) executes in a container without a GPU before box time is spent on them.

    python tools/gpu_tests_dryrun.py tests/test_zzz_buildindex.py -m gpu -q --timeout 300

A pass here says nothing about the CUDA path (the answers come from the CPU restatement); a failure is either a limitation of the
stand-in (entry points it does not model return an error code with the text "stub") or a genuine host-side bug of the test.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import _emu  # noqa: E402
import mauve_py_b200._capi as capi  # noqa: E402

stub = _emu.bench_stub_library()
capi.LIB_PATH = stub
os.environ["LD_PRELOAD"] = stub  # the C++ binaries the tests start (dropin_check, seam binaries)
import pytest  # noqa: E402

sys.exit(pytest.main(sys.argv[1:] + ["-p", "no:cacheprovider"]))
