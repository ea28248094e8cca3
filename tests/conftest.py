import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def built():
    """Build the host-side helper libraries once (synth generator, oracle)."""
    from spaghettisearch_b200 import _build
    _build.build_synth()
    from oracle import loader
    loader.build()
    return True


@pytest.fixture(scope="session")
def engine(built):
    """The CUDA engine through its C ABI; fails loudly when it cannot be created."""
    from spaghettisearch_b200 import capi
    e = capi.Engine(device=0, timing=True)
    yield e
    e.close()
