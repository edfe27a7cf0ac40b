import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib
    oracle_lib.lib()
    return oracle_lib


@pytest.fixture(scope="session")
def sky_inputs(oracle):
    """Earth atmosphere + oracle LUTs for the reference's default camera."""
    atmo = oracle.earth()
    cam = oracle.default_camera()
    trans, multi, view = oracle.sky_luts(atmo, cam.position[:])
    return atmo, trans, multi, view


@pytest.fixture(scope="session")
def blue_noise(oracle):
    return oracle.load_blue_noise()


@pytest.fixture()
def gpu_ctx():
    from minotert_b200 import capi
    ctx = capi.Context(0)  # raises (no fallback) when the library or the device is missing
    yield ctx
    ctx.close()
