import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (test infrastructure); built on demand with gcc."""
    from oracle import oracle as O

    O.build()
    O.lib()
    return O


@pytest.fixture(scope="session")
def sir_data():
    """inst/sir_incidence.csv of the reference, committed as a golden fixture
    (tests/golden/sir_incidence.csv; column `cases`)."""
    import numpy as np

    d = np.loadtxt(os.path.join(ROOT, "tests", "golden", "sir_incidence.csv"), delimiter=",",
                   skiprows=1)
    return {"day": d[:, 1].astype(int), "incidence": d[:, 0]}


def free_port():
    """A TCP port nobody listens on right now (torchrun rendezvous of the 2-rank tests): asked
    of the kernel, so two tests of one pytest process never meet on a port still in TIME_WAIT."""
    import socket

    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]
