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
    """Every GPU test runs once per kernel mapping: the one gmpe_create picks for the
    configuration ("auto": lane-per-agent gmpe_coop.cuh for N = L = 3 worlds, the tile kernel
    gmpe_step.cuh otherwise), the thread-per-world kernel (gmpe_small.cuh) and the tile kernel
    forced, so every mapping stays pinned to the oracle."""
    if metafunc.definition.get_closest_marker("gpu") is not None:
        metafunc.fixturenames.append("kernel_variant")
        metafunc.parametrize("kernel_variant", ["auto", "world", "tile"], indirect=True)


@pytest.fixture
def kernel_variant(request, monkeypatch):
    monkeypatch.delenv("GMPE_FORCE_TILE", raising=False)
    if request.param == "auto":
        monkeypatch.delenv("GMPE_KERNEL", raising=False)
    else:
        monkeypatch.setenv("GMPE_KERNEL", request.param)
    return request.param
