"""oracle/graph_oracle.py — TEST INFRASTRUCTURE ONLY.

NumPy restatement of the bulk graph build (vector-store_b200/csrc/graph_build.cu), i.e. of the
link-selection work `usearch::Index::add` does for the reference (call site
vs_index/usearch.rs:191-197).  Integer work on row indices, compared bit for bit with the GPU.
"""
from __future__ import annotations

import numpy as np

INVALID = np.uint32(0xFFFFFFFF)


def knn_lists(corpus, k_init: int, metric: int, storage: int, keys=None, alive=None):
    """Exact k_init-NN row indices of every row (self excluded), canonical order, ties by key."""
    from . import exact_topk
    n = corpus.shape[0]
    _, _, _, idx = exact_topk(corpus, corpus, min(k_init + 1, max(n, 1)), metric, storage, keys=keys, alive=alive)
    out = np.full((n, k_init), INVALID, dtype=np.uint32)
    for u in range(n):
        row = [v for v in idx[u] if v != INVALID and v != u][:k_init]
        out[u, :len(row)] = row
    return out


def prune_detour(knn: np.ndarray, R: int) -> np.ndarray:
    """edge u->L[j] gets detour[j] = #{i<j : L[j] in knn(L[i]) at position t<j}; keep R smallest (detour, j)."""
    n, k_init = knn.shape
    fwd = np.full((n, R), INVALID, dtype=np.uint32)
    for u in range(n):
        L = [int(v) for v in knn[u] if v != INVALID]
        rank = {v: j for j, v in enumerate(L)}
        det = [0] * len(L)
        for i, vi in enumerate(L):
            for t, y in enumerate(knn[vi]):
                if y == INVALID:
                    continue
                j = rank.get(int(y))
                if j is not None and j > i and t < j:
                    det[j] += 1
        order = sorted(range(len(L)), key=lambda j: (det[j], j))[:R]
        fwd[u, :len(order)] = [L[j] for j in order]
    return fwd


def reverse_edges(fwd: np.ndarray):
    """each node's reverse list = its R best proposers by (rank in the proposer's list, proposer index)."""
    n, R = fwd.shape
    props = [[] for _ in range(n)]
    for u in range(n):
        for r in range(R):
            v = fwd[u, r]
            if v != INVALID:
                props[int(v)].append((r, u))
    rev = np.full((n, R), INVALID, dtype=np.uint32)
    cnt = np.zeros(n, dtype=np.uint32)
    for v in range(n):
        p = sorted(props[v])[:R]
        rev[v, :len(p)] = [u for _, u in p]
        cnt[v] = len(p)
    return rev, cnt


def merge_graph(fwd: np.ndarray, rev: np.ndarray, cnt: np.ndarray, stride: int) -> np.ndarray:
    n, R = fwd.shape
    g = np.full((n, stride), INVALID, dtype=np.uint32)
    for u in range(n):
        out = []

        def push(v):
            if v == INVALID or v == u or len(out) >= R or int(v) in out:
                return
            out.append(int(v))
        for r in range(R // 2):
            push(fwd[u, r])
        for r in range(int(cnt[u])):
            push(rev[u, r])
        for r in range(R // 2, R):
            push(fwd[u, r])
        g[u, :len(out)] = out
    return g


def build_graph(corpus, k_init: int, R: int, metric: int, storage: int, keys=None, alive=None, stride=None):
    knn = knn_lists(corpus, k_init, metric, storage, keys=keys, alive=alive)
    if alive is not None:
        knn[np.asarray(alive) == 0] = INVALID
    fwd = prune_detour(knn, R)
    rev, cnt = reverse_edges(fwd)
    return merge_graph(fwd, rev, cnt, stride or ((R + 31) // 32 * 32))
