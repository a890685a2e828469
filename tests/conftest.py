import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def pytest_generate_tests(metafunc):
    """Every GPU test runs twice: with the kernel gmpe_create picks for the configuration
    (thread-per-world for N = L = 3 worlds, gmpe_small.cuh) and with the tile kernel forced
    (gmpe_step.cuh), so both mappings stay pinned to the oracle."""
    if metafunc.definition.get_closest_marker("gpu") is not None:
        metafunc.fixturenames.append("kernel_variant")
        metafunc.parametrize("kernel_variant", ["auto", "tile"], indirect=True)


@pytest.fixture
def kernel_variant(request, monkeypatch):
    if request.param == "tile":
        monkeypatch.setenv("GMPE_FORCE_TILE", "1")
    else:
        monkeypatch.delenv("GMPE_FORCE_TILE", raising=False)
    return request.param
