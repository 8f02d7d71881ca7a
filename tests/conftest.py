import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # The C-ABI library is built in-tree and git-ignored; on a fresh checkout build it once
    # (nvcc cross-compiles sm_100a without a GPU) so that the ABI / host-logic tests can load it.
    lib = os.path.join(ROOT, "metalquicha_b200", "libmqcb200.so")
    if not os.path.exists(lib):
        import __graft_entry__
        __graft_entry__.build()


@pytest.fixture(scope="session")
def engine():
    """One engine handle for the whole GPU session (the reference keeps one lazily
    created context per process, backends/cuest/backend/mqc_cuest_context.f90:285-299)."""
    from metalquicha_b200 import B200FockEngine
    eng = B200FockEngine(0)
    yield eng
    eng.close()
