import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def orc():
    import _oracle
    return _oracle.oracle_checker()


@pytest.fixture(scope="session")
def refc():
    import _oracle
    if not _oracle.have_ref():
        pytest.skip("oracle/_ref/libmauve_ref.so not built (needs /root/reference)")
    return _oracle.ref_checker()


@pytest.fixture(scope="session")
def mp():
    """the product package with the device initialised; GPU tests only"""
    import mauve_py_b200 as m
    from mauve_py_b200._capi import check
    check(m.lib().mcu_init(0))
    return m
