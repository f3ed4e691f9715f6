"""Host logic of the multi-GPU path on CPU: row-range sharding, the all-gather layout
[parts][q][k] and the (distance, key) merge, over world_size = 2 `gloo` processes.  Per-shard
results come from the CPU oracle here (tests only) — on a GPU box the same plumbing carries the
outputs of vsb_search_dev and the merge is the K8 CUDA kernel (tests/test_gpu_parity.py)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, dim, nq, k, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from importlib import import_module
    import oracle as O
    shard = import_module("vector_store_b200.host.shard")
    ds = import_module("vector_store_b200.host.datasets")
    x = ds.embedding_like(n, dim, n_clusters=8)
    q = ds.embedding_like(nq, dim, seed=4321, n_clusters=8)
    keys = (np.arange(n, dtype=np.uint64) * 7919) % np.uint64(1 << 40)
    lo, hi = shard.shard_range(n, rank, world)
    lk, ld, _, _ = O.exact_topk(x[lo:hi], q, k, O.COS, O.F32, keys=keys[lo:hi])
    gk, gd = shard.allgather_topk(torch.from_numpy(lk.view(np.int64)), torch.from_numpy(ld), world)
    mk, md = shard.merge_topk_host(gk.numpy().view(np.uint64), gd.numpy(), k)
    np.save(os.path.join(out_dir, f"k{rank}.npy"), mk)
    np.save(os.path.join(out_dir, f"d{rank}.npy"), md)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_partition_the_corpus():
    from importlib import import_module
    shard = import_module("vector_store_b200.host.shard")
    for n in (0, 1, 7, 1000, 1_000_003):
        for w in (1, 2, 3, 4, 8):
            r = [shard.shard_range(n, i, w) for i in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(180)
def test_two_rank_gloo_shard_allgather_merge_equals_global_exact(tmp_path):
    import oracle as O
    from importlib import import_module
    ds = import_module("vector_store_b200.host.datasets")
    n, dim, nq, k, world = 3001, 48, 33, 10, 2
    mp.spawn(_worker, args=(world, _free_port(), n, dim, nq, k, str(tmp_path)), nprocs=world, join=True)
    x = ds.embedding_like(n, dim, n_clusters=8)
    q = ds.embedding_like(nq, dim, seed=4321, n_clusters=8)
    keys = (np.arange(n, dtype=np.uint64) * 7919) % np.uint64(1 << 40)
    ok, od, _, _ = O.exact_topk(x, q, k, O.COS, O.F32, keys=keys)
    for r in range(world):  # every rank ends with the identical, globally exact result
        mk = np.load(tmp_path / f"k{r}.npy")
        md = np.load(tmp_path / f"d{r}.npy")
        assert np.array_equal(mk, ok)
        assert np.array_equal(md.view(np.uint32), od.view(np.uint32))


def _record_worker(rank, world, port, n, dim, nq, k, out_dir):
    """bench.py's exchange: ONE all-gather of a per-rank byte record [q*k keys (8 B) | q*k distances (4 B)], then a
    merge over parts that are `record` bytes apart (the strided K8 launch; restated with numpy views here)."""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from importlib import import_module
    import oracle as O
    shard = import_module("vector_store_b200.host.shard")
    ds = import_module("vector_store_b200.host.datasets")
    x = ds.embedding_like(n, dim, n_clusters=8)
    q = ds.embedding_like(nq, dim, seed=4321, n_clusters=8)
    keys = (np.arange(n, dtype=np.uint64) * 7919) % np.uint64(1 << 40)
    lo, hi = shard.shard_range(n, rank, world)
    lk, ld, _, _ = O.exact_topk(x[lo:hi], q, k, O.COS, O.F32, keys=keys[lo:hi])
    rec_bytes = nq * k * 12
    assert (nq * k) % 2 == 0                                    # keeps every rank's record 8-byte aligned (bench.py asserts it)
    rec_local = torch.empty(rec_bytes, dtype=torch.uint8)
    rec_local[:nq * k * 8].view(torch.int64).view(nq, k).copy_(torch.from_numpy(lk.view(np.int64)))
    rec_local[nq * k * 8:].view(torch.float32).view(nq, k).copy_(torch.from_numpy(ld))
    rec_all = torch.empty(world * rec_bytes, dtype=torch.uint8)
    dist.all_gather_into_tensor(rec_all, rec_local)
    raw = rec_all.numpy()
    gk = np.stack([raw[p * rec_bytes: p * rec_bytes + nq * k * 8].view(np.uint64).reshape(nq, k) for p in range(world)])
    gd = np.stack([raw[p * rec_bytes + nq * k * 8: (p + 1) * rec_bytes].view(np.float32).reshape(nq, k) for p in range(world)])
    mk, md = shard.merge_topk_host(gk, gd, k)
    np.save(os.path.join(out_dir, f"rk{rank}.npy"), mk)
    np.save(os.path.join(out_dir, f"rd{rank}.npy"), md)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_gloo_single_record_allgather_matches_global_exact(tmp_path):
    import oracle as O
    from importlib import import_module
    ds = import_module("vector_store_b200.host.datasets")
    n, dim, nq, k, world = 2503, 32, 24, 10, 2
    mp.spawn(_record_worker, args=(world, _free_port(), n, dim, nq, k, str(tmp_path)), nprocs=world, join=True)
    x = ds.embedding_like(n, dim, n_clusters=8)
    q = ds.embedding_like(nq, dim, seed=4321, n_clusters=8)
    keys = (np.arange(n, dtype=np.uint64) * 7919) % np.uint64(1 << 40)
    ok, od, _, _ = O.exact_topk(x, q, k, O.COS, O.F32, keys=keys)
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"rk{r}.npy"), ok)
        assert np.array_equal(np.load(tmp_path / f"rd{r}.npy").view(np.uint32), od.view(np.uint32))
