import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs /root/reference + oracle/_ref (build container only)")


def pytest_collection_modifyitems(config, items):
    from oracle import ref_loader
    have_ref = ref_loader.available()
    skip_ref = pytest.mark.skip(reason="compiled reference (oracle/_ref) not built")
    for item in items:
        if "reference" in item.keywords and not have_ref:
            item.add_marker(skip_ref)


@pytest.fixture(autouse=True)
def _device_buffer_guards(request):
    """With PP_GUARD=1 (GPU box; compute-sanitizer is closed on the pool) every device buffer of the library carries guard
    bytes; after each GPU test they are compared, so an out-of-bounds write fails the test that made it."""
    yield
    if os.environ.get("PP_GUARD", "0") not in ("", "0") and "gpu" in request.keywords:
        from protpore_b200 import _engine
        _engine.check_guards()
