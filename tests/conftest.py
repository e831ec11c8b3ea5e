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


@pytest.fixture(scope="session", autouse=True)
def _assembly_override():
    """MAGNETITE_B200_TEST_ASSEMBLY=1 runs every test that builds its options through the Python binding with the
    sorted-key assembly (mag_options.assembly = 1) unless the test sets the field itself: the whole parity suite
    against the non-default path with one command.  Unset (the default) nothing changes."""
    import os
    want = os.environ.get("MAGNETITE_B200_TEST_ASSEMBLY")
    if want is None:
        yield
        return
    from magnetite_b200 import _lib, solver
    original = _lib.default_options

    def patched(**overrides):
        overrides.setdefault("assembly", int(want))
        return original(**overrides)

    _lib.default_options = patched
    solver.default_options = patched
    try:
        yield
    finally:
        _lib.default_options = original
        solver.default_options = original
