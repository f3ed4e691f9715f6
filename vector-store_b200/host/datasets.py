"""Synthetic workloads of SURVEY §8d (seeds: corpus 1234, queries 4321)."""
from __future__ import annotations

import numpy as np


def sift_like(n: int, dim: int = 128, seed: int = 1234) -> np.ndarray:
    """C1: clip(round(|N(0,1)|*40), 0, 255) as f32 — integer valued, heavy-tailed norms, many exact ties."""
    rng = np.random.default_rng(seed)
    return np.clip(np.round(np.abs(rng.standard_normal((n, dim))) * 40.0), 0, 255).astype(np.float32)


def embedding_like(n: int, dim: int = 768, seed: int = 1234, n_clusters: int = 256, sigma: float = 0.3,
                   centers_seed: int = 99) -> np.ndarray:
    """C2/C3: Gaussian mixture (centres N(0,1), within-cluster sigma), L2-normalised."""
    crng = np.random.default_rng(centers_seed)
    centers = crng.standard_normal((n_clusters, dim)).astype(np.float32)
    rng = np.random.default_rng(seed)
    which = rng.integers(0, n_clusters, size=n)
    x = centers[which] + sigma * rng.standard_normal((n, dim), dtype=np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    return np.ascontiguousarray(x, dtype=np.float32)


# ---- N4: .fbin / .ibin files (crates/benchmark/src/data/fbin.rs:23-148: u32 count, u32 dimension, rows) ----
def read_bin_header(path: str) -> tuple[int, int]:
    with open(path, "rb") as f:
        count, dim = np.fromfile(f, dtype="<u4", count=2)
    return int(count), int(dim)


def read_fbin(path: str, start: int = 0, count: int | None = None) -> np.ndarray:
    n, dim = read_bin_header(path)
    count = n - start if count is None else min(count, n - start)
    return np.fromfile(path, dtype="<f4", count=count * dim, offset=8 + start * dim * 4).reshape(count, dim)


def read_ibin(path: str, start: int = 0, count: int | None = None) -> np.ndarray:
    n, dim = read_bin_header(path)
    count = n - start if count is None else min(count, n - start)
    return np.fromfile(path, dtype="<i4", count=count * dim, offset=8 + start * dim * 4).reshape(count, dim)


def write_fbin(path: str, rows: np.ndarray) -> None:
    rows = np.ascontiguousarray(rows, dtype="<f4")
    with open(path, "wb") as f:
        np.array(rows.shape, dtype="<u4").tofile(f)
        rows.tofile(f)


def write_ibin(path: str, rows: np.ndarray) -> None:
    rows = np.ascontiguousarray(rows, dtype="<i4")
    with open(path, "wb") as f:
        np.array(rows.shape, dtype="<u4").tofile(f)
        rows.tofile(f)
