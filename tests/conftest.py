import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))


def rel_err(a, b):
    """max |a-b| / max |b| - the 'relative' of north_star's 1e-4 (fp32) / 2e-2 (bf16) tolerances."""
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    if not torch.isfinite(a).all():
        return float("inf")
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


TOL_F32 = 1e-4     # north_star: activations and gradients within 1e-4 relative in fp32
TOL_BF16 = 2e-2    # ... and 2e-2 in bf16
