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
    from oracle import binding

    binding.build()
    binding.lib()
    return binding


@pytest.fixture(scope="session")
def pymodel():
    from oracle import pymodel as m

    return m


@pytest.fixture(scope="session")
def czk():
    import czk_b200

    return czk_b200


@pytest.fixture(scope="session")
def ctx(czk):
    """One GPU context for the whole session.  No fallback: without a device this raises."""
    c = czk.Context(int(os.environ.get("LOCAL_RANK", "0")))
    yield c
    c.close()
