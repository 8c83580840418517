import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


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
def oracle_lib():
    import oracle
    oracle.lib()
    return oracle


def make_cloud(rng, B, N, extent=25.0, duplicates=0):
    """Synthetic clouds of SURVEY 8(d): xyz ~ U(-extent, extent)^3; ``duplicates`` > 0 overwrites
    that many trailing points with copies of earlier ones (the reference pads short clouds with
    duplicated points, core/utils.py:103-106) to exercise exact ties."""
    pts = rng.uniform(-extent, extent, (B, N, 3)).astype(np.float32)
    if duplicates:
        for b in range(B):
            src = rng.randint(0, N - duplicates, duplicates)
            pts[b, N - duplicates:] = pts[b, src]
    return pts


def lattice_cloud(rng, B, N, step=0.5, side=12):
    """Points on a coarse integer lattice: masses of exactly equal distances (ties everywhere)."""
    g = rng.randint(0, side, (B, N, 3)).astype(np.float32) * np.float32(step)
    return g
