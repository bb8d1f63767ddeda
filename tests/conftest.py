import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from tests import oracle_binding

    return oracle_binding.load_oracle()


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference objects (oracle/_ref/libselas_ref.so); skipped when not built."""
    from tests import oracle_binding

    lib = oracle_binding.load_ref()
    if lib is None:
        pytest.skip("oracle/_ref/libselas_ref.so not built (reference tree absent)")
    return lib


@pytest.fixture(scope="session")
def ptexref():
    from tests import oracle_binding

    lib = oracle_binding.load_ptexref()
    if lib is None:
        pytest.skip("oracle/_ref/libptex_ref.so not built (reference tree absent)")
    return lib


@pytest.fixture(scope="session")
def ctx():
    """CUDA context of the product library; fails loudly (no fallback) when the library or the GPU is missing."""
    from selas_b200 import capi

    c = capi.Context(int(os.environ.get("SELAS_TEST_DEVICE", "0")))
    yield c
    c.close()
