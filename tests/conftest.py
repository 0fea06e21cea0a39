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
    """The product package.  The libraries are normally built by __graft_entry__.build(); a fresh checkout
    (the .so files are git-ignored) gets them compiled here once — compiling is not a compute fallback."""
    import gmu_water_simulation_b200 as pkg
    from gmu_water_simulation_b200 import binding

    if not (os.path.exists(binding._CUDA_SO) and os.path.exists(binding._HOST_SO)):
        pkg.build()
    return pkg


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle_binding

    oracle_binding.lib()
    return oracle_binding
