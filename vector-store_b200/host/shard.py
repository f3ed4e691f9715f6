"""Corpus sharding across ranks: every rank owns a contiguous row range of the corpus,
builds and searches only its shard, and the per-shard top-k lists are exchanged with ONE collective
(all-gather of [q, k] keys + distances, 12*k bytes per query per rank) and merged by (distance, key).

`torch.distributed` is plumbing only; the merge is the CUDA kernel behind vsb_merge_topk_dev.  The
CPU (gloo) path exists for the host-logic tests and takes the merge function from the caller."""
from __future__ import annotations

import numpy as np


def shard_range(n_rows: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous balanced split: the first n_rows % world shards get one extra row."""
    base, rem = divmod(n_rows, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def local_k(k: int, world: int) -> int:
    """Per-shard result length.  A query's true top-k is spread ~Binomial(k, 1/world) over the shards;
    k_local = k keeps the merged result exact for exact search and is the safe default for ANN."""
    return k


def allgather_topk(keys, dists, world: int):
    """keys: int64/uint64-as-int64 tensor [q, k], dists: f32 [q, k] -> ([world, q, k], [world, q, k])."""
    import torch
    import torch.distributed as dist
    gk = torch.empty((world,) + tuple(keys.shape), dtype=keys.dtype, device=keys.device)
    gd = torch.empty((world,) + tuple(dists.shape), dtype=dists.dtype, device=dists.device)
    if world == 1:
        gk[0].copy_(keys)
        gd[0].copy_(dists)
        return gk, gd
    # concatenated-along-dim-0 output form: accepted by both NCCL and gloo
    dist.all_gather_into_tensor(gk.view(-1, *keys.shape[1:]), keys.contiguous())
    dist.all_gather_into_tensor(gd.view(-1, *dists.shape[1:]), dists.contiguous())
    return gk, gd


def merge_topk_host(gk: np.ndarray, gd: np.ndarray, k: int):
    """NumPy statement of the K8 merge (tests only): [parts, q, k] -> [q, k] by (distance, key)."""
    parts, q, kk = gk.shape
    ak = gk.transpose(1, 0, 2).reshape(q, parts * kk)
    ad = gd.transpose(1, 0, 2).reshape(q, parts * kk)
    out_k = np.empty((q, k), dtype=gk.dtype)
    out_d = np.empty((q, k), dtype=gd.dtype)
    for i in range(q):
        o = np.lexsort((ak[i], ad[i]))[:k]
        out_k[i], out_d[i] = ak[i][o], ad[i][o]
    return out_k, out_d


# ---- cluster-routed sharding ---------------------------------------------------------------------------
# Row-range shards make every query visit every GPU, and beam-search cost is nearly flat in shard size,
# so QPS would not scale with the GPU count (SURVEY §8e caveat).  Routed sharding partitions the corpus
# by a coarse k-means quantiser instead: each shard owns whole clusters, and a query is searched only on
# the shards that own one of its `probes` nearest centroids; the other shards contribute empty lists to
# the same all-gather + merge.

def kmeans_centroids(sample: np.ndarray, n_centroids: int, assign_fn, iters: int = 8, seed: int = 7,
                     normalize: bool = True) -> np.ndarray:
    """Lloyd iterations; `assign_fn(centroids, rows) -> int array` is the GPU exact top-1 search."""
    rng = np.random.default_rng(seed)
    cent = sample[rng.choice(len(sample), n_centroids, replace=False)].astype(np.float32).copy()
    for _ in range(iters):
        a = assign_fn(cent, sample)
        sums = np.zeros_like(cent, dtype=np.float64)
        np.add.at(sums, a, sample)
        cnt = np.bincount(a, minlength=n_centroids)
        empty = cnt == 0
        cent = (sums / np.maximum(cnt, 1)[:, None]).astype(np.float32)
        if empty.any():  # re-seed empty clusters from random sample rows
            cent[empty] = sample[rng.choice(len(sample), int(empty.sum()), replace=False)]
        if normalize:
            cent /= np.maximum(np.linalg.norm(cent, axis=1, keepdims=True), 1e-20)
    return cent


def assign_owners(cluster_sizes: np.ndarray, world: int) -> np.ndarray:
    """Longest-processing-time bin packing of clusters onto shards (balances rows per GPU)."""
    owner = np.zeros(len(cluster_sizes), dtype=np.int64)
    load = np.zeros(world, dtype=np.int64)
    for c in np.argsort(-cluster_sizes, kind="stable"):
        r = int(np.argmin(load))
        owner[c] = r
        load[r] += int(cluster_sizes[c])
    return owner


def route_mask(nearest_centroids: np.ndarray, owner: np.ndarray, rank: int) -> np.ndarray:
    """nearest_centroids [q, probes] -> bool [q]: does shard `rank` own any of the query's probes."""
    return (owner[nearest_centroids] == rank).any(axis=1)
