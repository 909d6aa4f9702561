import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def ref():
    """compiled reference CPU path (oracle/_ref), test infrastructure only"""
    import refdirac
    if not refdirac.available():
        pytest.skip("oracle/_ref/libdirac_ref.so not built (make -C oracle)")
    return refdirac.load()


@pytest.fixture(scope="session")
def refser():
    """the compiled reference with the worker threads of its robust RTR / NSD solvers run
    synchronously (oracle/ref_shim_rtr_serial.c): deterministic pin of solver_mode 5 and 6"""
    import refdirac
    import os
    if not os.path.exists(refdirac.SERIAL_PATH):
        pytest.skip("oracle/_ref/libdirac_ref_serial.so not built (make -C oracle)")
    return refdirac.load_serial()


@pytest.fixture(scope="session")
def api():
    """the product library; GPU tests only"""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from sagecal_b200 import lib
    return lib.load()
