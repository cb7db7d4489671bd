import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

FIXTURES = os.path.join(ROOT, "tests", "fixtures")
GOLDEN = os.path.join(ROOT, "tests", "golden")
GENOME = os.path.join(FIXTURES, "electricsheep.247.11256.flam3")
VARIATIONS = os.path.join(FIXTURES, "variations.yaml")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _ensure_built():
    import refrakt_b200
    if not os.path.exists(refrakt_b200.LIB_PATH):
        from refrakt_b200 import build
        build.build()


@pytest.fixture(scope="session")
def rfk():
    _ensure_built()
    import refrakt_b200
    return refrakt_b200


@pytest.fixture(scope="session")
def compiler(rfk):
    return rfk.FlameCompiler(VARIATIONS)


@pytest.fixture(scope="session")
def flame(rfk, compiler):
    f = rfk.Flame.load_flame(GENOME, compiler)
    assert f is not None, rfk.Flame.last_error()
    return f


@pytest.fixture(scope="session")
def oracle_mod():
    import refrakt_oracle
    return refrakt_oracle


@pytest.fixture(scope="session")
def vt(oracle_mod):
    return oracle_mod.VariationTable(VARIATIONS)


@pytest.fixture(scope="session")
def oracle(oracle_mod, vt):
    return oracle_mod.Oracle(oracle_mod.load_flame(GENOME, vt), vt)


@pytest.fixture(scope="session")
def gpu_ready(rfk):
    """Fails (not skips) when the CUDA path cannot run: GPU tests must never pass on a fallback."""
    import ctypes
    rc = rfk.lib().rfk_set_device(0)
    assert rc == 0, "no CUDA device: " + rfk.lib().rfk_last_error().decode()
    return True
