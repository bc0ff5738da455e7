import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def po():
    """The CPU oracle (test infrastructure)."""
    from oracle import pyoracle
    pyoracle.lib()
    return pyoracle


@pytest.fixture(scope="session")
def b200():
    import astar_planners_b200  # noqa: F401
    import sys as _s
    return _s.modules["astar_planners_b200"]


@pytest.fixture(scope="session")
def synth():
    import astar_planners_b200  # noqa: F401
    from astar_planners_b200 import synth as s
    return s
