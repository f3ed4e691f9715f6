"""CPU checks of the error bounds the GPU path relies on (numpy emulation of the reduced-precision stages):
  * the exactness certificate of K3: |dot_tf32 - dot| <= (1.25 * 2^-9 + dim * 2^-22) |q||x|
  * the scaled-int8 traversal copy (VSB_FLAG_I8_TRAVERSAL): cosine error of the quantised rows
These mirror the constants in vector-store_b200/csrc/index.cu (exact_block) and convert.cu (convert_rows_i8s)."""
import numpy as np

from conftest import embedding_like, sift_like


def _tf32(a: np.ndarray, mode: str) -> np.ndarray:
    """fp32 -> TF32 operand (10 explicit mantissa bits): truncation or round-to-nearest-even."""
    u = a.astype(np.float32).view(np.uint32).astype(np.uint64)
    if mode == "rne":
        u = u + 0xFFF + ((u >> 13) & 1)
    return (u & 0xFFFFE000).astype(np.uint32).view(np.float32)


def test_tf32_dot_error_is_inside_the_certificate_bound():
    rng = np.random.default_rng(0)
    for make, dim in ((lambda n, s: embedding_like(n, 768, seed=s), 768), (lambda n, s: sift_like(n, 128, seed=s), 128),
                      (lambda n, s: rng.standard_normal((n, 100)).astype(np.float32) * 1e3, 100)):
        x, q = make(4000, 1), make(64, 2)
        exact = q.astype(np.float64) @ x.astype(np.float64).T
        scale = np.linalg.norm(q.astype(np.float64), axis=1)[:, None] * np.linalg.norm(x.astype(np.float64), axis=1)[None, :]
        rel_bound = 1.25 * 2.0 ** -9 + dim * 2.0 ** -22
        for mode in ("trunc", "rne"):
            approx = (_tf32(q, mode).astype(np.float32) @ _tf32(x, mode).astype(np.float32).T).astype(np.float64)
            err = np.abs(approx - exact) / scale
            assert err.max() < rel_bound, (mode, dim, err.max(), rel_bound)
            assert err.max() < 0.5 * rel_bound          # the bound is conservative by design (Cauchy-Schwarz)


def test_fp32_summation_order_error_is_inside_the_certificate_bound():
    # two different fp32 summation orders of the same products differ by far less than dim * 2^-22 |q||x|
    x, q = embedding_like(2000, 768, seed=3), embedding_like(32, 768, seed=4)
    prod = q[:, None, :] * x[None, :, :]
    fwd = np.zeros(prod.shape[:2], np.float32)
    rev = np.zeros(prod.shape[:2], np.float32)
    for j in range(768):
        fwd += prod[:, :, j]
        rev += prod[:, :, 767 - j]
    scale = np.linalg.norm(q, axis=1)[:, None] * np.linalg.norm(x, axis=1)[None, :]
    assert (np.abs(fwd - rev) / scale).max() < 768 * 2.0 ** -22


def _quantize_i8_scaled(a: np.ndarray):
    """numpy restatement of convert_rows_i8s_kernel: per-row scale max|x|/127, q = rint(x/scale)."""
    mx = np.abs(a).max(axis=1, keepdims=True)
    scale = np.where(mx > 0, mx / 127.0, 1.0).astype(np.float32)
    qv = np.clip(np.rint(a / scale), -127, 127).astype(np.int32)
    nrm_units = np.linalg.norm(a.astype(np.float64), axis=1) / scale[:, 0]
    return qv, nrm_units


def test_scaled_int8_cosine_error_is_small_against_neighbour_gaps():
    x, q = embedding_like(20000, 768, seed=5), embedding_like(50, 768, seed=6)
    xq, xn = _quantize_i8_scaled(x)
    qq, qn = _quantize_i8_scaled(q)
    approx = 1.0 - (qq.astype(np.float64) @ xq.astype(np.float64).T) / (qn[:, None] * xn[None, :])
    exact = 1.0 - (q.astype(np.float64) @ x.astype(np.float64).T)        # unit vectors
    err = np.abs(approx - exact)
    assert err.max() < 4e-3 and err.mean() < 5e-4    # measured: max 2.2e-3 over 10^6 pairs, mean 3.1e-4
    # what matters for the traversal: the true top-10 stay inside the int8 top-40 (the 4k re-rank window)
    top_true = np.argsort(exact, axis=1)[:, :10]
    top_i8 = np.argsort(approx, axis=1)[:, :40]
    kept = np.mean([len(np.intersect1d(a, b)) / 10 for a, b in zip(top_true, top_i8)])
    assert kept >= 0.999


def test_sampled_list_bound_leaves_enough_rows():
    """The all-pairs build bounds every kNN list by the m-th best row of a tile-strided sample (exact_block /
    tc_sample_threshold_kernel): sample = 32 tiles of 256 rows spread over the block, m = 3 k' sample / rows.
    Emulated in NumPy on the bench's corpus generator (131 072 rows = the all-pairs prefix of every large build): the bound
    must leave at least k = 65 rows (k_init + self) for all but a small fraction of the queries — the library redoes a block
    without bounds above 2 % — and on average ~3 k' of them."""
    from importlib import import_module
    ds = import_module("vector_store_b200.host.datasets")
    rows, dim, kp, k = 131_072, 64, 96, 65
    x = ds.embedding_mix(rows, dim, seed=1234, n_clusters=2560)
    q = x[::257][:400]                                      # queries are rows of the block (all-pairs)
    d = 1.0 - q.astype(np.float64) @ x.astype(np.float64).T  # cosine distance of unit rows
    tile_step = rows // 256 // 32 * 256
    cols = np.concatenate([np.arange(t * tile_step, t * tile_step + 256) for t in range(32)])
    m = int(min(32, max(4, np.ceil(3.0 * kp * len(cols) / rows))))
    assert m == 18
    bound = np.sort(d[:, cols], axis=1)[:, m - 1]
    below = (d <= bound[:, None]).sum(axis=1)
    short = float((below < k).mean())
    print(f"sampled bound: mean {below.mean():.0f} rows below it (3 k' = {3 * kp}), min {below.min()}, short lists {short:.4f}")
    assert short <= 0.02
    assert 1.5 * kp < below.mean() < 6 * kp
    # rows stored in the order of their clusters break the sample (whole tiles of one cluster): that is what the
    # library's short-list count + redo is for — the emulation must SEE the failure mode it guards against
    order = np.argsort((x @ x[0]).astype(np.float64), kind="stable")
    ds_sorted = d[:, order]
    bound_s = np.sort(ds_sorted[:, cols], axis=1)[:, m - 1]
    below_s = (ds_sorted <= bound_s[:, None]).sum(axis=1)
    assert below_s.std() > below.std()                      # far more erratic than on shuffled rows


def test_mma_fragment_mapping_of_the_tensor_core_traversal():
    """group_reduce_mma (graph_search.cuh) fills mma.sync.m16n8k16 fragments straight from the SIMT kernel's loads: lane
    l = 4 g + t holds 16 contiguous bytes (8 elements) of row u, of row u + 1 and of the query; elements 0..3 feed one
    MMA, 4..7 the next.  With the PTX fragment layouts (A: a0 = (g, 2t..), a1 = (g+8, 2t..), a2 = (g, 2t+8..),
    a3 = (g+8, 2t+8..); B: b0 = (2t.., n = g), b1 = (2t+8.., n = g); D: c0,c1 = (g, 2t..), c2,c3 = (g+8, 2t..)) the
    diagonal D[g][g] is quad g's partial dot product for row u and D[g+8][g] for row u + 1, held by thread t = g / 2 in
    register g % 2.  Emulated here with exact integer arithmetic."""
    rng = np.random.default_rng(3)
    segs = 3                                   # 768 bf16 elements = 3 warp loads of 512 bytes
    row_a = rng.integers(-8, 9, size=(segs, 32, 8)).astype(np.int64)   # [segment][lane][element]
    row_b = rng.integers(-8, 9, size=(segs, 32, 8)).astype(np.int64)
    qry = rng.integers(-8, 9, size=(segs, 32, 8)).astype(np.int64)
    D = np.zeros((16, 8), dtype=np.int64)
    for s in range(segs):
        for half in (0, 4):                    # first MMA: elements 0..3, second: 4..7
            A = np.zeros((16, 16), dtype=np.int64)
            B = np.zeros((16, 8), dtype=np.int64)
            for lane in range(32):
                g, t = lane // 4, lane % 4
                A[g, 2 * t:2 * t + 2] = row_a[s, lane, half:half + 2]               # a0
                A[g + 8, 2 * t:2 * t + 2] = row_b[s, lane, half:half + 2]           # a1
                A[g, 2 * t + 8:2 * t + 10] = row_a[s, lane, half + 2:half + 4]      # a2
                A[g + 8, 2 * t + 8:2 * t + 10] = row_b[s, lane, half + 2:half + 4]  # a3
                B[2 * t:2 * t + 2, g] = qry[s, lane, half:half + 2]                 # b0
                B[2 * t + 8:2 * t + 10, g] = qry[s, lane, half + 2:half + 4]        # b1
            D += A @ B
    v0 = v1 = 0
    for lane in range(32):
        g, t = lane // 4, lane % 4
        if t == g // 2:                        # the thread that holds the diagonal entry of quad g
            c = [D[g, 2 * t], D[g, 2 * t + 1], D[g + 8, 2 * t], D[g + 8, 2 * t + 1]]
            assert 2 * t + (g % 2) == g
            v0 += c[g % 2]
            v1 += c[2 + g % 2]
    assert v0 == int((row_a * qry).sum()) and v1 == int((row_b * qry).sum())
