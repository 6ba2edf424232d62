import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import binding
    return binding


@pytest.fixture(scope="session")
def refso(oracle):
    if oracle.ref is None:
        pytest.skip("oracle/_ref/libslamref.so not built (needs /root/reference)")
    # pin the function-static Shift_Amount of the reference's area estimator (quirk Q10)
    oracle.ref.ref_init_area_shift(0.01, 0.05)
    return oracle.ref


@pytest.fixture(scope="session")
def sg():
    """the product package (ctypes over lib/libslamgpu.so); no fallback if it is missing"""
    import slam_constructor_b200 as pkg
    pkg.lib()
    return pkg


@pytest.fixture(scope="session")
def gpu(sg):
    ctx = sg.Context(0)
    yield ctx
    ctx.close()
