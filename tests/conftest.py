import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")
    config.addinivalue_line("markers", "slow: takes more than ~20 s")


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference build (oracle/_ref) -- checker only."""
    import oracle
    if not os.path.exists(oracle.REFERENCE_SO):
        pytest.skip("oracle/_ref/libiwref.so not built (make -C oracle ref needs /root/reference)")
    return oracle.load_reference()


@pytest.fixture(scope="session")
def emu():
    """TEST-ONLY host emulation of the kernel bodies: lets the host logic (plan, weight lists, tables, API) be
    exercised on a box without a GPU.  Built into tests/_emu; never part of the product."""
    from imageworsener_b200 import _build, libimagew as L
    return L.Lib(_build.build_emulator())


@pytest.fixture(scope="session")
def prod():
    """The product: the CUDA library, through its C ABI.  No fallback of any kind."""
    import torch
    from imageworsener_b200 import libimagew as L
    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device")
    return L.load_product()


@pytest.fixture(scope="session")
def oracle_port():
    from imageworsener_b200 import _build
    so = os.path.join(ROOT, "oracle", "libiworacle.so")
    if not os.path.exists(so):
        _build.build_oracle()
    import ctypes
    return ctypes.CDLL(so)
