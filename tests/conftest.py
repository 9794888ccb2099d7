import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    # a GPU test on a box without a GPU must FAIL LOUDLY when explicitly selected with -m gpu,
    # and is simply deselected by -m "not gpu"; nothing to do here.
    return


@pytest.fixture(scope="session")
def golden():
    return {n[:-4]: np.load(os.path.join(GOLDEN, n), allow_pickle=False)
            for n in os.listdir(GOLDEN) if n.endswith(".npz")}


def expand_posterior(lab, alt, w1, w2, V):
    """Same compact → dense posterior expansion as oracle/make_golden.py (kept in sync by test)."""
    import torch
    B, T = lab.shape
    base = (1.0 - w1 - w2) / V
    p = np.repeat(base[..., None], V, axis=2).astype(np.float32)
    bi, ti = np.meshgrid(np.arange(B), np.arange(T), indexing="ij")
    p[bi, ti, lab] += w1
    p[bi, ti, alt] += w2
    return torch.from_numpy(p)
