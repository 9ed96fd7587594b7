import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def kats():
    with open(os.path.join(ROOT, "tests", "golden", "reference_kats.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def oracle():
    """The C oracle (oracle/jwave_oracle.c); builds it on first use."""
    from oracle import c_oracle
    c_oracle.lib()
    return c_oracle


@pytest.fixture(scope="session")
def jw():
    import jwave_pro_b200
    return jwave_pro_b200


@pytest.fixture(scope="session")
def gpu_ctx(jw):
    """One CUDA context for the whole GPU session.  Fails (never skips) when the native path is unavailable."""
    return jw.default_context()


from jwave_pro_b200.synth import chirp, splitmix_uniform  # noqa: E402,F401
