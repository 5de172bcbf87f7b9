import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def gl_ref() -> bytes:
    with open(os.path.join(GOLDEN, "GL000193.1_p1.ref"), "rb") as f:
        return f.read()


@pytest.fixture(scope="session")
def gl_fm(gl_ref):
    """The oracle's rebuild of the shipped GL000193.1_p1.fm (byte-identical; see test_oracle_golden.py)."""
    from oracle import bwt_oracle as O
    return O.build_fm(gl_ref, 1)


@pytest.fixture(scope="session")
def gl_coracle(gl_fm, gl_ref):
    from oracle.c_oracle import COracle
    return COracle(gl_fm.L, gl_fm.SA, gl_fm.pointers, gl_fm.shift, gl_ref, mode=0)


@pytest.fixture(scope="session")
def gl_coracle_cf(gl_fm, gl_ref):
    from oracle.c_oracle import COracle
    return COracle(gl_fm.L, gl_fm.SA, gl_fm.pointers, gl_fm.shift, gl_ref, mode=1)
