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


# ---- counter-based mixture generator: NumPy twin of tools/synth/synth.cu (bit for bit) ----
_K_A, _K_B, _K_CLUSTER = np.uint64(0x9E3779B97F4A7C15), np.uint64(0xD1B54A32D192ED03), np.uint64(0xC1057E7)
_INV_STD = np.float32(float.fromhex("0x1.bb67aep-16"))  # 1 / sqrt(4 * (65536^2 - 1) / 12)


def _mix64(z: np.ndarray) -> np.ndarray:
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def _h3(seed, a, b) -> np.ndarray:
    return _mix64(np.uint64(seed) + a * _K_A + b * _K_B)


def _gauss(h: np.ndarray) -> np.ndarray:
    m = np.uint64(0xFFFF)
    s = ((h & m) + ((h >> np.uint64(16)) & m) + ((h >> np.uint64(32)) & m) + (h >> np.uint64(48))).astype(np.int64) - 131070
    return s.astype(np.float32) * _INV_STD


def embedding_mix(n: int, dim: int = 768, row0: int = 0, seed: int = 1234, n_clusters: int = 256, sigma: float = 0.3,
                  centers_seed: int = 99) -> np.ndarray:
    """Rows [row0, row0 + n) of the counter-based Gaussian-mixture corpus, L2-normalised — exactly what
    tools/synth/synth.cu writes into HBM for the same arguments (element (r, j) depends only on (seed, r, j)):
    integer hashing for the randomness, individually rounded fp32 operations in the kernel's order for the rest."""
    out = np.empty((n, dim), dtype=np.float32)
    with np.errstate(over="ignore"):
        j1 = np.arange(1, dim + 1, dtype=np.uint64)[None, :]
        centers = _gauss(_h3(centers_seed, np.arange(n_clusters, dtype=np.uint64)[:, None], j1))
        lanes = np.arange(32)
        sig = np.float32(sigma)
        CH = 8192

        def fill(c0: int) -> None:
            r = np.arange(row0 + c0, row0 + min(c0 + CH, n), dtype=np.uint64)
            which = (_h3(np.uint64(seed) ^ _K_CLUSTER, r, np.uint64(0)) % np.uint64(n_clusters)).astype(np.int64)
            x = centers[which] + sig * _gauss(_h3(seed, r[:, None], j1))
            acc = np.zeros((len(r), 32), dtype=np.float32)
            for t in range(0, dim, 32):
                blk = x[:, t:t + 32]
                acc[:, :blk.shape[1]] = acc[:, :blk.shape[1]] + blk * blk
            for s in (16, 8, 4, 2, 1):
                acc = acc + acc[:, lanes ^ s]
            inv = np.float32(1.0) / np.sqrt(acc[:, 0])
            out[c0:c0 + len(r)] = x * inv[:, None]

        starts = list(range(0, n, CH))
        if len(starts) > 1:  # NumPy releases the GIL inside its loops: chunks run on all host cores
            import os
            from concurrent.futures import ThreadPoolExecutor
            with ThreadPoolExecutor(max_workers=os.cpu_count() or 1) as pool:
                list(pool.map(fill, starts))
        else:
            for c0 in starts:
                fill(c0)
    return out


_synth = None


def synth_lib():
    """tools/synth/libvsbsynth.so (built by __graft_entry__.build()); bench / test tooling, not the product library."""
    global _synth
    if _synth is None:
        import ctypes as C
        import os
        path = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tools", "synth",
                            "libvsbsynth.so")
        lib = C.CDLL(path)
        lib.vsbsynth_embedding_mix.restype = C.c_int
        lib.vsbsynth_embedding_mix.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_float,
                                               C.c_uint64, C.c_uint64, C.c_void_p]
        _synth = lib
    return _synth


def embedding_mix_dev(d_out: int, n: int, dim: int, row0: int = 0, seed: int = 1234, n_clusters: int = 256,
                      sigma: float = 0.3, centers_seed: int = 99, stream: int = 0) -> None:
    """Same rows, written to the device buffer `d_out` (raw pointer, n x dim f32) by the CUDA generator."""
    rc = synth_lib().vsbsynth_embedding_mix(d_out, row0, n, dim, n_clusters, sigma, seed, centers_seed, stream or None)
    if rc != 0:
        raise RuntimeError(f"vsbsynth_embedding_mix failed ({rc})")


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


# ---- N4: parquet datasets (crates/benchmark/src/data/parquet.rs: the VectorDBBench directory layout) ----
# <dir>/*train*.parquet  : columns `id` (int64), `emb` (list<float32|float64>)          -> corpus rows
# <dir>/test.parquet      : columns `id`, `emb`                                           -> queries
# <dir>/neighbors.parquet : columns `id`, `neighbors_id` (list<int64>)                    -> ground truth per query id
class ParquetConfig:
    """Same fields and defaults as the reference's parquet `Config` (data/parquet.rs:34-103)."""

    def __init__(self, ext: str = "parquet", train_file_pattern: str = "train", test_file_name: str = "test.parquet",
                 neighbors_file_name: str = "neighbors.parquet", id_column: str = "id", embedding_column: str = "emb",
                 neighbors_id_column: str = "neighbors_id"):
        self.ext = ext
        self.train_file_pattern = train_file_pattern
        self.test_file_name = test_file_name
        self.neighbors_file_name = neighbors_file_name
        self.id_column = id_column
        self.embedding_column = embedding_column
        self.neighbors_id_column = neighbors_id_column


def _pq():
    import pyarrow.parquet as pq  # imported lazily: only dataset tooling needs pyarrow
    return pq


def _embeddings_to_f32(col) -> np.ndarray:
    """list<float32|float64> (or large_list) column chunk -> [rows][dim] f32 (f64 is narrowed, parquet.rs:151-169)."""
    import pyarrow as pa
    col = col.combine_chunks() if isinstance(col, pa.ChunkedArray) else col
    if len(col) == 0:
        return np.empty((0, 0), np.float32)
    if col.null_count:
        raise ValueError("null embedding")
    values = col.flatten().to_numpy(zero_copy_only=False)
    offsets = col.offsets.to_numpy()
    dims = np.diff(offsets)
    if not np.all(dims == dims[0]):
        raise ValueError("ragged embeddings in one batch")
    return np.ascontiguousarray(values.reshape(len(col), int(dims[0])), dtype=np.float32)


def parquet_train_files(path: str, config: ParquetConfig | None = None) -> list[str]:
    """Regular files in `path` whose name contains the train pattern and whose extension matches (parquet.rs:124-149)."""
    import os
    cfg = config or ParquetConfig()
    out = []
    for name in sorted(os.listdir(path)):
        full = os.path.join(path, name)
        if os.path.isfile(full) and cfg.train_file_pattern in name and name.rsplit(".", 1)[-1] == cfg.ext and "." in name:
            out.append(full)
    return out


def parquet_dimension(path: str, config: ParquetConfig | None = None) -> int:
    """Length of the first test embedding (parquet.rs:105-123)."""
    import os
    cfg = config or ParquetConfig()
    f = _pq().ParquetFile(os.path.join(path, cfg.test_file_name))
    batch = next(f.iter_batches(batch_size=1, columns=[cfg.embedding_column]))
    return int(_embeddings_to_f32(batch.column(0)).shape[1])


def parquet_vector_batches(path: str, config: ParquetConfig | None = None):
    """Yields (ids int64 [n], rows f32 [n][dim]) per row group of every train file (parquet.rs:229-330): the
    unit a caller hands to vsb_add."""
    cfg = config or ParquetConfig()
    for file in parquet_train_files(path, cfg):
        f = _pq().ParquetFile(file)
        for rg in range(f.num_row_groups):
            t = f.read_row_group(rg, columns=[cfg.id_column, cfg.embedding_column])
            if t.num_rows == 0:
                continue
            ids = t.column(cfg.id_column).combine_chunks().to_numpy(zero_copy_only=False).astype(np.int64)
            yield ids, _embeddings_to_f32(t.column(cfg.embedding_column))


def parquet_queries(path: str, config: ParquetConfig | None = None, id_ok=None, limit: int = 10):
    """-> list of (query f32 [dim], set of ground-truth ids): test rows joined with neighbors.parquet by id; each
    neighbour list is filtered by `id_ok`, cut to its first `limit` survivors and dropped when empty
    (parquet.rs:332-433)."""
    import os
    cfg = config or ParquetConfig()
    pq = _pq()
    t = pq.read_table(os.path.join(path, cfg.test_file_name), columns=[cfg.id_column, cfg.embedding_column])
    q_ids = t.column(cfg.id_column).combine_chunks().to_numpy(zero_copy_only=False).astype(np.int64)
    q_rows = _embeddings_to_f32(t.column(cfg.embedding_column))
    queries = {int(i): q_rows[n] for n, i in enumerate(q_ids)}
    nt = pq.read_table(os.path.join(path, cfg.neighbors_file_name), columns=[cfg.id_column, cfg.neighbors_id_column])
    n_ids = nt.column(cfg.id_column).combine_chunks().to_numpy(zero_copy_only=False).astype(np.int64)
    neigh = nt.column(cfg.neighbors_id_column).combine_chunks()
    offsets = neigh.offsets.to_numpy()
    values = neigh.flatten().to_numpy(zero_copy_only=False).astype(np.int64)
    truth = {}
    for n, i in enumerate(n_ids):
        ids = values[offsets[n]:offsets[n + 1]]
        kept = [int(v) for v in ids if id_ok is None or id_ok(int(v))][:limit]
        if kept:
            truth[int(i)] = set(kept)
    return [(queries[i], truth[i]) for i in queries if i in truth]


def write_parquet_dataset(path: str, ids: np.ndarray, rows: np.ndarray, q_ids: np.ndarray, queries: np.ndarray,
                          neighbors: np.ndarray, train_files: int = 1, row_group_rows: int | None = None,
                          emb_type: str = "float32") -> None:
    """Writes a dataset in the layout above (test fixture / export of a synthetic corpus)."""
    import os
    import pyarrow as pa
    pq = _pq()
    os.makedirs(path, exist_ok=True)
    ft = pa.float32() if emb_type == "float32" else pa.float64()

    def emb_col(m):
        flat = pa.array(np.ascontiguousarray(m).reshape(-1).astype(np.float32 if emb_type == "float32" else np.float64), ft)
        offs = pa.array(np.arange(0, (len(m) + 1) * m.shape[1], m.shape[1], dtype=np.int64))
        return pa.LargeListArray.from_arrays(offs, flat)

    per = (len(ids) + train_files - 1) // train_files
    for f in range(train_files):
        sl = slice(f * per, min(len(ids), (f + 1) * per))
        tab = pa.table({"id": pa.array(ids[sl].astype(np.int64)), "emb": emb_col(rows[sl])})
        name = "train.parquet" if train_files == 1 else f"shuffle_train-{f:02d}-of-{train_files:02d}.parquet"
        pq.write_table(tab, os.path.join(path, name), row_group_size=row_group_rows or max(1, sl.stop - sl.start))
    pq.write_table(pa.table({"id": pa.array(q_ids.astype(np.int64)), "emb": emb_col(queries)}),
                   os.path.join(path, "test.parquet"))
    noffs = pa.array(np.arange(0, (len(neighbors) + 1) * neighbors.shape[1], neighbors.shape[1], dtype=np.int64))
    ncol = pa.LargeListArray.from_arrays(noffs, pa.array(neighbors.reshape(-1).astype(np.int64)))
    pq.write_table(pa.table({"id": pa.array(q_ids.astype(np.int64)), "neighbors_id": ncol}),
                   os.path.join(path, "neighbors.parquet"))
