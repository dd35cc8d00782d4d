import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "nvidia-jetson-workload_b200"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _gpu_count():
    try:
        from weather_sim import _capi
        return _capi.device_count()
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    """`-m gpu` on a box without a usable CUDA device (or without the built extension) must FAIL loudly, not pass
    by skipping: the product has no CPU fallback. Every selected gpu test gets a failing setup in that case."""
    markexpr = (config.getoption("markexpr", "") or "").replace(" ", "")
    if markexpr != "gpu" or _gpu_count() > 0:
        return
    for item in items:
        if item.get_closest_marker("gpu") is not None:
            item.fixturenames.insert(0, "_wsb_no_gpu")


@pytest.fixture
def _wsb_no_gpu():
    pytest.fail("pytest -m gpu needs a CUDA device and lib/libweather_b200.so (the product has no CPU fallback): "
                "0 devices visible", pytrace=False)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session", autouse=True)
def _build_oracle():
    """The oracle is the checker for every test: make sure its C library exists."""
    import oracle_py
    oracle_py.build_oracle()
    yield
