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
