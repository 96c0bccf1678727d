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
    # no test may hang a (GPU) box: pytest-timeout kills anything beyond 10 minutes (multi-process tests set tighter limits)
    if config.pluginmanager.hasplugin("timeout"):
        for item in items:
            if item.get_closest_marker("timeout") is None:
                item.add_marker(pytest.mark.timeout(600))
    # the multi-GPU test spawns its own ranks: under an outer torchrun every rank would spawn on the same devices / port
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        for item in items:
            if "test_multigpu_nccl" in item.nodeid:
                item.add_marker(pytest.mark.skip(reason="run this file with plain pytest, not under torch.distributed.run"))
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
