import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the native pieces once per session (no-op when up to date; prebuilt .so files travel to the GPU box)."""
    import __graft_entry__ as g

    try:
        g.build_tools()
        g.build_oracle()
    except Exception:
        if not (os.path.exists(os.path.join(ROOT, "oracle", "libisonclust2_oracle.so")) and
                os.path.exists(os.path.join(ROOT, "isonclust2_b200", "libicb_synth.so"))):
            raise
    if not os.path.exists(g.LIB):
        g.build_cuda()
    try:
        g.build_cli()
    except Exception:
        if not os.path.exists(os.path.join(ROOT, "isonclust2_b200", "bin", "isONclust2")):
            raise


@pytest.fixture(scope="session")
def port():
    import oracle
    return oracle.port()


@pytest.fixture(scope="session")
def ref():
    import oracle
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    return oracle.ref()


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_v1.npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def ctx():
    from isonclust2_b200 import api
    c = api.Context(0)
    yield c
    c.close()
