import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def built():
    """Make sure both shared libraries exist (no-op when up to date)."""
    import __graft_entry__ as g
    from magnetite_b200 import _lib
    from oracle import oracle as O
    if not _lib.LIB_PATH.exists():
        g.build()
    O.build()
    return True


@pytest.fixture(scope="session")
def ctx(built):
    from magnetite_b200 import _lib
    c = _lib.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def material():
    from magnetite_b200 import meshgen
    return meshgen.EXAMPLE_MATERIAL
