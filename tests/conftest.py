import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built_library():
    """The CUDA extension must exist; build it in-tree if a fresh checkout has not done so."""
    from graphax_b200 import runtime
    if not runtime.LIB_PATH.exists():
        from graphax_b200.build import build_library
        build_library()
    return runtime.load_library()
