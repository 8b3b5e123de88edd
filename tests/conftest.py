import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built():
    """Compile the test infrastructure (oracle restatement, reference object when its sources are
    here, host-side emulation double).  The product .so is built by __graft_entry__.build()."""
    from bensolve_b200 import build
    build.build_oracle()
    return build


@pytest.fixture(scope="session")
def oracle_lib(built):
    from bensolve_b200 import capi
    return capi.load_lib(capi.ORACLE_SO)


@pytest.fixture(scope="session")
def ref_lib(built):
    from bensolve_b200 import capi
    if not os.path.exists(capi.REF_SO):
        pytest.skip("oracle/_ref/libref_poly.so not built (reference sources absent)")
    return capi.load_lib(capi.REF_SO)


@pytest.fixture(scope="session")
def emul_lib(built):
    from bensolve_b200 import capi
    return capi.load_lib(built.build_emulation())


@pytest.fixture(scope="session")
def product_lib():
    """The CUDA engine.  No fallback: if it is not built or no GPU is visible the test fails."""
    from bensolve_b200 import capi
    lib = capi.load_product()
    assert lib.b200_device_count() > 0, "no CUDA device visible"
    return lib
