import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
if os.path.dirname(__file__) not in sys.path:
    sys.path.insert(0, os.path.dirname(__file__))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.oracle_py import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def golden():
    import json
    import numpy as np
    here = os.path.join(os.path.dirname(__file__), "golden")
    with open(os.path.join(here, "known_answers.json")) as f:
        answers = json.load(f)
    vec = np.load(os.path.join(here, "vectors.npz"))
    return answers, vec


@pytest.fixture(scope="session")
def pkg():
    import __graft_entry__ as g
    if not os.path.exists(os.path.join(g.PKG_DIR, "libsahbvh_b200.so")):
        g.build()
    return g.load_pkg()


@pytest.fixture(scope="session", params=["binary", "wide"])
def emul(request):
    """Host build of the traversal core over the binary flattened tree and over the 4-wide
    quantised tree (csrc/wide.c): every host-logic test runs against both."""
    from common import HostEmul
    return HostEmul(request.param)


@pytest.fixture(scope="session")
def reflib():
    from oracle.oracle_py import RefLib, ref_available
    if not ref_available():
        pytest.skip("compiled reference (oracle/_ref/libref.so) not available")
    return RefLib()
