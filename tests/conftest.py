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


@pytest.fixture(scope="session")
def sprites():
    d = np.load(os.path.join(GOLDEN, "sprites.npz"))
    return {"front": d["front"], "right": d["right"]}


@pytest.fixture(scope="session")
def palette_golden():
    return dict(np.load(os.path.join(GOLDEN, "palette_golden.npz")))


@pytest.fixture(scope="session")
def hist_golden():
    return dict(np.load(os.path.join(GOLDEN, "hist_golden.npz")))


@pytest.fixture(scope="session")
def reference_run():
    """Outputs of the reference's own source files executed over oracle/ref_shim.py (oracle/run_reference.py)."""
    return dict(np.load(os.path.join(GOLDEN, "reference_run.npz")))


@pytest.fixture(scope="session")
def cuda():
    import torch

    if not torch.cuda.is_available():
        pytest.fail("a test marked gpu ran without a CUDA device — there is no CPU fallback to exercise")
    return torch.device("cuda:0")


def sprite_like_batch(rng, batch, hw=64, min_colors=16, max_colors=48, opaque=0.165):
    """Synthetic RGBA sprites with the statistics of the reference dataset: a few dozen
    colours, ~16.5 % opaque pixels, transparent pixels black; uint8 (B,hw,hw,4)."""
    out = np.zeros((batch, hw, hw, 4), np.uint8)
    for b in range(batch):
        n = int(rng.integers(min_colors, max_colors + 1))
        cols = rng.integers(0, 256, size=(n, 4), dtype=np.int64).astype(np.uint8)
        cols[:, 3] = 255
        mask = rng.random((hw, hw)) < opaque
        pick = rng.integers(0, n, size=(hw, hw))
        out[b][mask] = cols[pick[mask]]
    return out
