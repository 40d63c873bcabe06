import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden", "reference_vectors.npz")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    return dict(np.load(GOLDEN))


def state_dict_from(golden, prefix):
    import torch
    tag = prefix + "_sd::"
    return {k[len(tag):]: torch.from_numpy(v.copy()) for k, v in golden.items() if k.startswith(tag)}


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


SHIPPED = os.path.join(ROOT, "tests", "golden", "shipped_w43_weights.npz")


@pytest.fixture(scope="session")
def shipped():
    return dict(np.load(SHIPPED))


def shipped_state_dict(shipped, tag):
    import torch
    pre = tag + "::"
    return {k[len(pre):]: torch.from_numpy(v.copy()) for k, v in shipped.items() if k.startswith(pre)}
