import os
import sys

import pytest

os.environ.setdefault("PTB_POISON", "1")      # pt_launch_host starts frame 0 from NaNs: an unwritten pixel cannot pass a comparison

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def port():
    import oracle_api
    return oracle_api.load_port()


@pytest.fixture(scope="session")
def ref():
    """The reference's own programs compiled for the host (oracle/_ref); skip when it cannot exist."""
    import oracle_api
    lib = oracle_api.load_ref()
    if lib is None:
        pytest.skip("oracle/_ref/libptref.so not built (no reference checkout on this machine)")
    return lib


@pytest.fixture(scope="session")
def cornell():
    from acgpathtracing_b200 import scenes
    return scenes.cornell32()


@pytest.fixture(scope="session")
def material_scene():
    from acgpathtracing_b200 import scenes
    return scenes.material_scene(2)
