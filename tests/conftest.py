import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built_lib():
    """libeuclider_b200.so, built in-tree if stale (nvcc cross-compiles without a GPU)."""
    from euclider_b200 import _build

    _build.build()
    from euclider_b200 import lib

    return lib()


@pytest.fixture(scope="session")
def oracle():
    import oracle_api

    oracle_api.lib()
    return oracle_api
