import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def oracle():
    """CPU oracle (test infrastructure; oracle/og_oracle.c)."""
    from oracle import ogoracle
    ogoracle.lib()
    return ogoracle


@pytest.fixture(scope="session")
def ctx():
    """GPU context through the C ABI; fails loudly if the CUDA library is missing."""
    from obsgrid_b200.api import Context
    c = Context(int(os.environ.get("LOCAL_RANK", "0")))
    yield c
    c.close()
