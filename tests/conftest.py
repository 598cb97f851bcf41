"""pytest configuration: the `gpu` marker and shared fixtures.

`-m "not gpu"` runs in the GPU-less build container (oracle vs golden vectors, host logic, C-ABI
export check, gloo world-size-2 tests).  `-m gpu` runs on a B200 through the C ABI and needs
neither /root/reference nor the network.
"""
import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)
GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def golden():
    return load_golden


@pytest.fixture(scope="session")
def cuda_lib():
    """Builds (if needed) and loads librecsys_b200.so; fails loudly when that is impossible."""
    from recommender_b200 import build, _lib
    build.build()
    _lib.lib.load()
    return _lib.lib


@pytest.fixture
def fake_kernels(monkeypatch):
    """Host-logic tests on torch-CPU: the C-ABI calls of layers / model / optimizers are served by tests/fake_ops.py (numpy
    oracle + torch-CPU).  The product itself has no CPU path."""
    import tests.fake_ops as fake
    import recommender_b200.layers as layers
    import recommender_b200.model as model
    import recommender_b200.optimizers as optimizers
    for mod in (layers, model, optimizers):
        monkeypatch.setattr(mod, "ops", fake)
    return fake
