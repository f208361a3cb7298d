import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _native_built():
    """Build what can be built here: the product .so (nvcc cross-compiles without a GPU), the
    test-only host-logic library, and - where /root/reference exists - the compiled reference."""
    import __graft_entry__ as g
    import shutil
    # On a GPU box the snapshot's prebuilt artefacts are what is under test (and what the driver records as loaded): never
    # spend GPU time recompiling there.  In the build container everything is (re)built from the sources.
    if os.path.exists("/dev/nvidiactl"):
        yield
        return
    if shutil.which("nvcc"):
        g.build_cuda()
    g.build_io()
    g.build_hostlogic()
    try:
        g.build_oracle()
    except Exception:
        pass
    yield


def have_ref():
    from oracle import ref
    return ref.available()


needs_ref = pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libsvref.so"))
                               and not os.path.isdir("/root/reference"),
                               reason="compiled reference (oracle/_ref) not available")
