import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def toy():
    with open(os.path.join(GOLDEN, "toy_golden.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def layer_golden():
    return dict(np.load(os.path.join(GOLDEN, "layer_golden.npz")))


def synthetic_kg(V, R, E, seed=1234, skewed=False):
    """SURVEY.md 8(d) generator: uniform or skewed (s,o = floor(V*u^3) relabelled, r = floor(R*u^2))."""
    rng = np.random.Generator(np.random.PCG64(seed))
    if not skewed:
        s = rng.integers(0, V, E)
        o = rng.integers(0, V, E)
        r = rng.integers(0, R, E)
    else:
        perm = rng.permutation(V)
        s = perm[np.minimum((V * rng.random(E) ** 3).astype(np.int64), V - 1)]
        o = perm[np.minimum((V * rng.random(E) ** 3).astype(np.int64), V - 1)]
        r = np.minimum((R * rng.random(E) ** 2).astype(np.int64), R - 1)
    return np.stack([s, r, o], 1).astype(np.int32)
