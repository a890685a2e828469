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
    configuration ("auto": lane-per-agent gmpe_coop.cuh for worlds with 2 <= N = L <= 10 and no
    obstacles / walls, the tile kernel gmpe_step.cuh otherwise) and the tile kernel forced
    (GMPE_FLAG_FORCE_TILE), so both mappings stay pinned to the oracle."""
    if metafunc.definition.get_closest_marker("gpu") is not None:
        metafunc.fixturenames.append("kernel_variant")
        metafunc.parametrize("kernel_variant", ["auto", "tile"], indirect=True)


@pytest.fixture
def kernel_variant(request, monkeypatch):
    from gmpe_b200 import config

    monkeypatch.setattr(config, "FORCE_TILE_DEFAULT", request.param == "tile")
    return request.param
