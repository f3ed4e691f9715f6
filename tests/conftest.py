import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def _has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


# ---- synthetic data -------------------------------------------------------------------
def sift_like(n, dim=128, seed=1234):
    """C1: clip(round(|N(0,1)|*40), 0, 255) as f32 — integer valued, many exact distance ties."""
    rng = np.random.default_rng(seed)
    return np.clip(np.round(np.abs(rng.standard_normal((n, dim))) * 40.0), 0, 255).astype(np.float32)


def embedding_like(n, dim=768, seed=1234, n_clusters=256, sigma=0.3, centers_seed=99):
    """C2: Gaussian mixture (centres N(0,1), within-cluster sigma), L2-normalised."""
    crng = np.random.default_rng(centers_seed)
    centers = crng.standard_normal((n_clusters, dim)).astype(np.float32)
    rng = np.random.default_rng(seed)
    which = rng.integers(0, n_clusters, size=n)
    x = centers[which] + sigma * rng.standard_normal((n, dim)).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    return x.astype(np.float32)
