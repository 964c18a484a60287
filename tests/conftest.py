import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

GOLDEN = os.path.join(ROOT, "tests", "golden")


# The parity tests pin the list build to the two-pass kernels that every committed GPU result was taken with.  The one-pass cell
# build, the library's auto mode (its default: it times both builds and self-checks the first one-pass build against a two-pass
# build) and the other result-neutral variants have their own tests in test_zzzzzzz_variants.py, collected last, so that a
# problem in a variant cannot mask the parity results of everything else.  Tests that exercise a variant set the variable themselves.
os.environ.setdefault("DDCB200_LISTBUILD", "twopass")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
