"""oracle — CPU restatement of the reference's index path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import this package; the product (``vector-store_b200``) never does.

Parity status:
  * exact distances / exact top-k: PINNED by the reference's golden vectors G1-G10
    (tests/test_oracle_golden.py), restated from the `usearch` 2.22.0 crate's published semantics
    (Cargo.toml:93; the crate itself is not buildable offline) and the in-tree wrappers
    vs_index/usearch.rs:142-251,442-513,1179-1205, distance.rs:58-105, similarity.rs:26-37.
  * ANN recall vs real USearch: PARITY UNPINNED — `HnswCpu` is a USearch-equivalent HNSW
    restatement, not USearch; recall claims are always "vs exact ground truth".
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_DIR, "_build")

L2SQ, COS, IP, HAMMING = 0, 1, 2, 3
F32, F16, BF16, I8, B1 = 0, 1, 2, 3, 4
METRICS = {"l2sq": L2SQ, "euclidean": L2SQ, "cos": COS, "cosine": COS, "ip": IP, "dot": IP, "dotproduct": IP,
           "hamming": HAMMING}
SCALARS = {"f32": F32, "f16": F16, "bf16": BF16, "i8": I8, "b1": B1}


def build(force: bool = False) -> None:
    """Compiles oracle/exact.c and oracle/hnsw_cpu.cpp into oracle/_build/ (gcc/g++, seconds)."""
    need = force or not all(os.path.exists(os.path.join(_BUILD, f)) for f in ("liboracle_exact.so", "libhnsw_cpu.so"))
    if not need:
        srcs = [os.path.join(_DIR, f) for f in ("exact.c", "hnsw_cpu.cpp", "Makefile")]
        outs = [os.path.join(_BUILD, f) for f in ("liboracle_exact.so", "libhnsw_cpu.so")]
        need = max(os.path.getmtime(s) for s in srcs) > min(os.path.getmtime(o) for o in outs)
    if need:
        subprocess.run(["make", "-C", _DIR, "-s"] + (["-B"] if force else []), check=True)


_exact = None
_hnsw = None


def _lib_exact():
    global _exact
    if _exact is None:
        build()
        lib = C.CDLL(os.path.join(_BUILD, "liboracle_exact.so"))
        lib.vso_row_bytes.restype = C.c_uint32
        lib.vso_row_bytes.argtypes = [C.c_int, C.c_uint32]
        lib.vso_convert_row.argtypes = [C.c_int, C.c_void_p, C.c_uint32, C.c_void_p]
        lib.vso_sqnorm.restype = C.c_float
        lib.vso_sqnorm.argtypes = [C.c_int, C.c_void_p, C.c_uint32]
        lib.vso_exact_topk.argtypes = [C.c_int, C.c_int, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64,
                                       C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p]
        lib.vso_distance_matrix.argtypes = [C.c_int, C.c_int, C.c_uint32, C.c_void_p, C.c_uint64, C.c_void_p,
                                            C.c_uint64, C.c_void_p]
        lib.vso_f32_to_b1x8.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
        _exact = lib
    return _exact


def _lib_hnsw():
    global _hnsw
    if _hnsw is None:
        build()
        lib = C.CDLL(os.path.join(_BUILD, "libhnsw_cpu.so"))
        lib.hnsw_create.restype = C.c_void_p
        lib.hnsw_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_uint64]
        lib.hnsw_free.argtypes = [C.c_void_p]
        lib.hnsw_set_ef.argtypes = [C.c_void_p, C.c_int]
        lib.hnsw_size.restype = C.c_uint64
        lib.hnsw_size.argtypes = [C.c_void_p]
        lib.hnsw_add_batch.restype = C.c_int
        lib.hnsw_add_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int]
        lib.hnsw_search_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p,
                                          C.c_int, C.c_int]
        lib.hnsw_search_one.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        lib.hnsw_max_threads.restype = C.c_int
        _hnsw = lib
    return _hnsw


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def row_bytes(storage: int, dim: int) -> int:
    return int(_lib_exact().vso_row_bytes(storage, dim))


def convert_rows(rows, storage: int) -> np.ndarray:
    """f32 rows -> padded storage rows (uint8 [n, row_bytes]); the cast applied on add and on the query."""
    rows = _f32(np.atleast_2d(rows))
    n, dim = rows.shape
    rb = row_bytes(storage, dim)
    out = np.zeros((n, rb), dtype=np.uint8)
    lib = _lib_exact()
    for i in range(n):
        lib.vso_convert_row(storage, rows[i].ctypes.data_as(C.c_void_p), dim, out[i].ctypes.data_as(C.c_void_p))
    return out


def exact_topk(corpus, queries, k: int, metric: int, storage: int = F32, keys=None, alive=None):
    """Exact k-NN in the canonical fp32 order; ties by key.  Returns (keys, dists, counts, row_idx)."""
    corpus = _f32(np.atleast_2d(corpus))
    queries = _f32(np.atleast_2d(queries))
    n, dim = corpus.shape if corpus.size else (0, queries.shape[1])
    nq = queries.shape[0]
    keys_a = None if keys is None else np.ascontiguousarray(keys, dtype=np.uint64)
    alive_a = None if alive is None else np.ascontiguousarray(alive, dtype=np.uint8)
    ok = np.empty((nq, k), dtype=np.uint64)
    od = np.empty((nq, k), dtype=np.float32)
    oc = np.empty(nq, dtype=np.uint32)
    oi = np.empty((nq, k), dtype=np.uint32)
    _lib_exact().vso_exact_topk(storage, metric, dim, _ptr(corpus), _ptr(keys_a), _ptr(alive_a), n, _ptr(queries), nq,
                                k, _ptr(ok), _ptr(od), _ptr(oc), _ptr(oi))
    return ok, od, oc, oi


def distance_matrix(corpus, queries, metric: int, storage: int = F32) -> np.ndarray:
    corpus = _f32(np.atleast_2d(corpus))
    queries = _f32(np.atleast_2d(queries))
    out = np.empty((queries.shape[0], corpus.shape[0]), dtype=np.float32)
    _lib_exact().vso_distance_matrix(storage, metric, corpus.shape[1], _ptr(corpus), corpus.shape[0], _ptr(queries),
                                     queries.shape[0], _ptr(out))
    return out


def exact_topk_f64(corpus, queries, k: int, metric: int):
    """Independent float64 NumPy formulation (f32 storage only) used to cross-check exact.c."""
    x = np.asarray(corpus, dtype=np.float64)
    q = np.asarray(queries, dtype=np.float64)
    if metric == L2SQ:
        d = ((q[:, None, :] - x[None, :, :]) ** 2).sum(-1)
    else:
        dot = q @ x.T
        if metric == IP:
            d = 1.0 - dot
        else:
            qn = np.linalg.norm(q, axis=1)[:, None]
            xn = np.linalg.norm(x, axis=1)[None, :]
            with np.errstate(divide="ignore", invalid="ignore"):
                d = 1.0 - dot / (qn * xn)
            both = (qn == 0) & (xn == 0)
            one = ((qn == 0) | (xn == 0)) & ~both
            d = np.where(both, 0.0, np.where(one, 1.0, d))
            d = np.clip(d, 0.0, 2.0)
    idx = np.argsort(d, axis=1, kind="stable")[:, :k]
    return idx, np.take_along_axis(d, idx, axis=1)


def f32_to_b1x8(v) -> np.ndarray:
    """vs_index/usearch.rs:1179-1205."""
    v = _f32(v).ravel()
    out = np.zeros((v.size + 7) // 8, dtype=np.uint8)
    _lib_exact().vso_f32_to_b1x8(_ptr(v), v.size, _ptr(out))
    return out


# ---- A9 / A10: output contract the engine's results must satisfy ------------------------------------------
def distance_is_valid(value: float, metric: int, dim: int | None = None) -> bool:
    """distance.rs:58-105 `Distance::try_from((f32, SpaceType, Option<Dimensions>))`."""
    v = float(np.float32(value))
    if metric == COS:
        return 0.0 <= v <= 2.0
    if metric == L2SQ:
        return v >= 0.0  # NaN fails, +inf passes
    if metric == IP:
        return not np.isnan(v)
    if not (v >= 0.0) or not np.isfinite(v) or v != np.floor(v) or dim is None:
        return False
    return v <= float(dim)


def similarity_score(value: float, metric: int, dim: int | None = None) -> np.float32:
    """similarity.rs:26-37 `SimilarityScore::from(Distance)` (f32 arithmetic)."""
    d = np.float32(value)
    if metric in (COS, IP):
        return np.float32((np.float32(2.0) - d) / np.float32(2.0))
    if metric == L2SQ:
        return np.float32(np.float32(1.0) / (np.float32(1.0) + d))
    return np.float32(np.float32(1.0) - d / np.float32(dim))


def recall_at_k(found_keys: np.ndarray, true_keys: np.ndarray) -> float:
    """|returned ∩ true_top_k| / min(k, |GT|)  (latte/vector-search/metrics.rn:24-40; benchmark/src/db.rs:308)."""
    hits = 0
    total = 0
    invalid = np.uint64(0xFFFFFFFFFFFFFFFF)
    for f, t in zip(found_keys, true_keys):
        t = t[t != invalid]
        total += len(t)
        hits += len(np.intersect1d(f[f != invalid], t, assume_unique=False))
    return hits / max(total, 1)


class HnswCpu:
    """USearch-equivalent CPU HNSW (restatement, not USearch 2.22.0) — see hnsw_cpu.cpp."""

    def __init__(self, dim: int, metric: int, capacity: int, connectivity: int = 16, expansion_add: int = 128,
                 expansion_search: int = 64, storage: int = F32, seed: int = 42, threads: int = 0):
        if metric == HAMMING or storage not in (F32, BF16):
            raise ValueError("hnsw_cpu supports f32/bf16 storage with l2sq/cos/ip")
        self._lib = _lib_hnsw()
        self._h = self._lib.hnsw_create(dim, metric, connectivity, expansion_add, expansion_search, capacity, seed)
        self.dim, self.storage = dim, storage
        self.threads = threads or (os.cpu_count() or 1)

    def add(self, keys, rows) -> None:
        rows = _f32(np.atleast_2d(rows))
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        rc = self._lib.hnsw_add_batch(self._h, _ptr(keys), _ptr(rows), rows.shape[0], self.storage, self.threads)
        if rc != 0:
            raise RuntimeError("hnsw_cpu: capacity exceeded")

    def set_ef(self, ef: int) -> None:
        self._lib.hnsw_set_ef(self._h, ef)

    def search(self, queries, k: int):
        queries = _f32(np.atleast_2d(queries))
        ok = np.empty((queries.shape[0], k), dtype=np.uint64)
        od = np.empty((queries.shape[0], k), dtype=np.float32)
        self._lib.hnsw_search_batch(self._h, _ptr(queries), queries.shape[0], k, _ptr(ok), _ptr(od), self.storage,
                                    self.threads)
        return ok, od

    def search_one(self, query, k: int):
        query = _f32(query).ravel()
        ok = np.empty(k, dtype=np.uint64)
        od = np.empty(k, dtype=np.float32)
        self._lib.hnsw_search_one(self._h, _ptr(query), k, _ptr(ok), _ptr(od))
        return ok, od

    def __len__(self) -> int:
        return int(self._lib.hnsw_size(self._h))

    def close(self) -> None:
        if self._h:
            self._lib.hnsw_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
