import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


def pytest_sessionstart(session):
    """The suite exercises the built library and CLI: build them when they are missing (nvcc cross-compiles
    without a GPU).  Staleness is not checked here: snapshots do not preserve modification times."""
    from centrifuge_b200 import build as b
    if not (os.path.exists(b.LIB) and os.path.exists(b.CLI)):
        b.build(verbose=False)


@pytest.fixture(scope="session")
def adv_base():
    import util
    return util.golden_index("adv")


@pytest.fixture(scope="session")
def adv_reads():
    import lzma
    import util
    p = os.path.join(util.CACHE, "golden", "adv.reads.fa")
    os.makedirs(os.path.dirname(p), exist_ok=True)
    with lzma.open(os.path.join(util.GOLDEN, "adv.reads.fa.xz")) as f, open(p, "wb") as g:
        g.write(f.read())
    return p
