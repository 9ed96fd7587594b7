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


def splitmix_uniform(seed, shape):
    """Counter-based uniform(-1, 1) stream shared by tests and bench."""
    n = int(np.prod(shape))
    z = (np.arange(1, n + 1, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)) + np.uint64(seed)
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    z = z ^ (z >> np.uint64(31))
    u = (z >> np.uint64(11)).astype(np.float64) * (2.0 ** -53)
    return (u * 2.0 - 1.0).reshape(shape)


def chirp(batch, n):
    t = np.arange(n, dtype=np.float64) / n
    f0 = 4.0 + (np.arange(batch) % 13)[:, None]
    k = n / 8.0
    return np.sin(2.0 * np.pi * (f0 * t[None, :] + 0.5 * k * t[None, :] ** 2))
