import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden_dir():
    return ROOT / "tests" / "golden"


@pytest.fixture(autouse=True)
def _fresh_shadows(request):
    """Diagnostic switch (CREAM_TEST_CLEAR_SHADOWS=1): drop every cached bf16 weight shadow before each GPU test."""
    import os
    if os.environ.get("CREAM_TEST_CLEAR_SHADOWS") == "1" and request.node.get_closest_marker("gpu") is not None:
        from cream_b200 import ops
        ops.SHADOWS.clear()
    yield
