import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def gws():
    """The product package; libraries must already be built (no CPU fallback, no JIT at import)."""
    import gmu_water_simulation_b200 as pkg

    return pkg


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle_binding

    oracle_binding.lib()
    return oracle_binding
