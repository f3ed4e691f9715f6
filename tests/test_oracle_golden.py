"""Pins the CPU oracle against every golden vector / known-answer test the reference holds for the
index path.  Paths cited are relative to the reference repository."""
import numpy as np
import pytest

import oracle as O
from oracle import graph_oracle


def nearest(corpus, keys, q, k, metric=O.L2SQ, storage=O.F32, alive=None):
    kk, dd, cc, _ = O.exact_topk(np.array(corpus, np.float32), np.array([q], np.float32), k, metric, storage,
                                 keys=np.array(keys, np.uint64), alive=alive)
    return [int(x) for x in kk[0][:cc[0]]], [float(x) for x in dd[0][:cc[0]]]


def test_g1_add_or_replace_size_ann():
    # crates/vector-store/src/vs_index/usearch.rs:1298-1458
    corpus = [[1, 1, 1], [2, -2, 2], [3, 3, 3]]
    q = [2.2, -2.2, 2.2]
    assert nearest(corpus, [1, 2, 3], q, 1)[0] == [2]
    corpus2 = [[1, 1, 1], [2, -2, 2], [2.1, -2.1, 2.1]]  # remove(3) + add(3, ...)
    assert nearest(corpus2, [1, 2, 3], q, 1)[0] == [3]
    assert nearest(corpus2, [1, 2, 3], q, 1, alive=np.array([1, 1, 0], np.uint8))[0] == [2]


def test_g2_similarity_scores_are_decreasing_and_correctly_converted():
    # crates/vector-store/tests/integration/vs_index.rs:1746-1887
    keys, dists = nearest([[0.0], [1.0], [3.0]], [10, 11, 13], [0.0], 3)
    assert dists == [0.0, 1.0, 9.0]
    scores = [float(O.similarity_score(d, O.L2SQ)) for d in dists]
    assert np.allclose(scores, [1.0, 0.5, 0.1], atol=1e-5)
    assert scores[0] > scores[1] > scores[2]


def test_g3_simple_l2_nearest():
    # tests/integration/vs_index.rs:226-311
    assert nearest([[1, 1, 1], [2, -2, 2], [3, 3, 3]], [1, 2, 3], [2.1, -2, 2], 1)[0] == [2]


def test_g4_quantization_is_effectively_applied():
    # tests/integration/quantization.rs:23-124: F32 < 0.1 ; I8 > 300 (= 507)
    _, d32 = nearest([[0.9, 0.1, 0.1]], [1], [1.0, 0.0, 0.0], 1, storage=O.F32)
    _, d8 = nearest([[0.9, 0.1, 0.1]], [1], [1.0, 0.0, 0.0], 1, storage=O.I8)
    assert d32[0] < 0.1
    assert d8[0] > 300 and d8[0] == 507.0
    rows = O.convert_rows([[0.9, 0.1, 0.1]], O.I8)
    assert list(rows[0, :3].view(np.int8)) == [114, 13, 13]


@pytest.mark.parametrize("storage", [O.F32, O.F16, O.BF16, O.I8, O.B1])
def test_g5_self_distance_is_exactly_zero(storage):
    # tests/integration/quantization.rs:175-259 (dim 1536, all 0.5)
    x = np.full((1, 1536), 0.5, np.float32)
    _, d = nearest(x, [7], x[0], 1, storage=storage)
    assert d == [0.0]


def test_g6_b1_dimension_not_multiple_of_8():
    # tests/integration/quantization.rs:293-358 (dim 100)
    x = np.linspace(-1, 1, 100, dtype=np.float32)[None, :]
    _, d = nearest(x, [1], x[0], 1, storage=O.B1)
    assert d == [0.0]
    _, d = nearest(x, [1], -x[0], 1, storage=O.B1)
    assert d == [100.0 - 0.0] or d == [float((x[0] > 0).sum() + (-x[0] > 0).sum())]


def test_g7_f32_to_b1x8():
    # vs_index/usearch.rs:1622-1664
    assert O.f32_to_b1x8([]).size == 0
    assert list(O.f32_to_b1x8([1, 1, 1, 1, 0, 0, 0, 0])) == [0b00001111]
    assert list(O.f32_to_b1x8([1, 0, 1, 0, 1, 0, 1, 0, -1, -1, -1, -1, 1, 1, 1, 1])) == [0b01010101, 0b11110000]
    assert list(O.f32_to_b1x8([1.0] * 64)) == [0xFF] * 8
    assert list(O.f32_to_b1x8([1, 0, 1, 0, 1, 0, 1, 0, 1, -1, 1])) == [0b01010101, 0b00000101]


def test_g8_distance_validation_tables():
    # distance.rs:123-195
    inf, nan = float("inf"), float("nan")
    ok_e = [0.0, 0.123, 1.0, 2.0, 5.0, 100.5, 3.4e38, inf]
    bad_e = [-0.1, -1.0, -inf, nan]
    assert all(O.distance_is_valid(v, O.L2SQ) for v in ok_e) and not any(O.distance_is_valid(v, O.L2SQ) for v in bad_e)
    ok_c = [0.0, 0.123, 1.0, 2.0]
    bad_c = [5.0, 100.5, 3.4e38, -0.1, -1.0, inf, -inf, nan]
    assert all(O.distance_is_valid(v, O.COS) for v in ok_c) and not any(O.distance_is_valid(v, O.COS) for v in bad_c)
    ok_d = [0.0, 0.123, 1.0, 2.0, 5.0, 100.5, 3.4e38, -0.1, -1.0, inf, -inf]
    assert all(O.distance_is_valid(v, O.IP) for v in ok_d) and not O.distance_is_valid(nan, O.IP)
    ok_h = [0.0, 1.0, 2.0]
    bad_h = [0.123, 5.0, 100.5, 3.4e38, -0.1, -1.0, inf, -inf, nan]
    assert all(O.distance_is_valid(v, O.HAMMING, 3) for v in ok_h)
    assert not any(O.distance_is_valid(v, O.HAMMING, 3) for v in bad_h)


def test_g8_similarity_exact_values():
    # similarity.rs:46-132
    f = lambda v, m, d=None: float(O.similarity_score(v, m, d))
    assert f(0.0, O.L2SQ) == 1.0 and f(1.0, O.L2SQ) == 0.5 and f(99.0, O.L2SQ) == np.float32(0.01)
    assert f(1000.0, O.L2SQ) < 0.001
    assert f(0.0, O.COS) == 1.0 and f(1.0, O.COS) == 0.5 and f(2.0, O.COS) == 0.0
    assert f(6.7, O.IP) == np.float32(-2.35) and f(-1.8, O.IP) == np.float32(1.9)
    assert f(64.0, O.HAMMING, 128) == 0.5 and f(35.0, O.HAMMING, 50) == np.float32(0.3)
    assert f(128.0, O.HAMMING, 128) == 0.0


def test_g9_similarity_function_semantics():
    # crates/validator/src/similarity_functions.rs:113-178: axis vectors + a 4th, query (1,0,0)
    base = [[1, 0, 0], [0, 1, 0], [0, 0, 1]]
    q = [1, 0, 0]
    k, d = nearest(base + [[2, 0, 0]], [1, 2, 3, 4], q, 4, O.L2SQ)
    assert k[0] == 1 and d[0] == 0.0 and d[1] > 0.0                    # EUCLIDEAN => {1}
    k, d = nearest(base + [[2, 0, 0]], [1, 2, 3, 4], q, 4, O.COS)
    assert set(k[:2]) == {1, 4} and d[0] == d[1] == 0.0 and d[2] > 0    # COSINE => {1, 4}
    k, d = nearest(base + [[2, 0, 0]], [1, 2, 3, 4], q, 4, O.IP)
    assert k[0] == 4 and d[0] == -1.0                                   # DOT_PRODUCT => {4}


def test_g10_empty_index_returns_empty():
    # tests/integration/vs_index.rs:1889-1951
    kk, dd, cc, _ = O.exact_topk(np.zeros((0, 3), np.float32), np.zeros((1, 3), np.float32), 5, O.L2SQ)
    assert cc[0] == 0 and np.all(kk == np.uint64(0xFFFFFFFFFFFFFFFF)) and np.all(np.isinf(dd))


def test_g12_f32_cosine_resolution():
    # crates/validator/src/quantization_and_rescoring.rs:21-37, 98-154: 500 vectors q + i*1e-3*(2,4,8)
    q = np.array([1.0, 1.0, 1.0], np.float32)
    x = np.stack([q + i * 1e-3 * np.array([2, 4, 8], np.float32) for i in range(500)]).astype(np.float32)
    idx64, _ = O.exact_topk_f64(x, q[None, :], 500, O.COS)
    kk, dd, _, _ = O.exact_topk(x, q[None, :], 500, O.COS, keys=np.arange(500))
    assert np.all(np.diff(dd[0]) >= 0)
    # fp32 resolves the ordering except among the first few near-identical rows (distance < 1e-6)
    assert np.mean(kk[0].astype(np.int64) == idx64[0]) > 0.95
    assert list(kk[0][50:]) == list(range(50, 500))


@pytest.mark.parametrize("metric", [O.L2SQ, O.COS, O.IP])
def test_exact_c_matches_independent_float64(metric):
    rng = np.random.default_rng(7)
    x = rng.standard_normal((3000, 96)).astype(np.float32)
    q = rng.standard_normal((16, 96)).astype(np.float32)
    kk, dd, _, ii = O.exact_topk(x, q, 10, metric)
    i64, d64 = O.exact_topk_f64(x, q, 10, metric)
    assert np.array_equal(ii, i64.astype(np.uint32))
    assert np.allclose(dd, d64, rtol=1e-5, atol=2e-6)


def test_ties_are_broken_by_key():
    x = np.zeros((6, 4), np.float32)
    keys = np.array([50, 10, 40, 20, 60, 30], np.uint64)
    kk, dd, cc, _ = O.exact_topk(x, np.zeros((1, 4), np.float32), 4, O.L2SQ, keys=keys)
    assert list(kk[0]) == [10, 20, 30, 40] and cc[0] == 4


def test_storage_casts():
    v = np.array([[1.0, 1.0 + 2 ** -8, 1.0 + 3 * 2 ** -8, -0.3, 1e-3, 70000.0, 0.0, 2.5]], np.float32)
    bf = O.convert_rows(v, O.BF16)[0, :16].view(np.uint16)
    # exact halfway cases round to the even mantissa
    assert bf[0] == 0x3F80 and bf[1] == 0x3F80 and bf[2] == 0x3F82
    f16 = O.convert_rows(v, O.F16)[0, :16].view(np.float16)
    assert np.array_equal(f16, v[0].astype(np.float16))
    i8 = O.convert_rows(np.array([[0.5, -0.5, 2.0, -2.0, 0.004]], np.float32), O.I8)[0, :5].view(np.int8)
    assert list(i8) == [64, -64, 127, -127, 1]
    assert O.row_bytes(O.F32, 3) == 16 and O.row_bytes(O.BF16, 768) == 1536 and O.row_bytes(O.B1, 100) == 16


def test_graph_oracle_small():
    rng = np.random.default_rng(3)
    x = rng.standard_normal((200, 16)).astype(np.float32)
    g = graph_oracle.build_graph(x, k_init=16, R=8, metric=O.L2SQ, storage=O.F32)
    assert g.shape == (200, 32)
    deg = (g != graph_oracle.INVALID).sum(1)
    assert deg.max() <= 8 and deg.min() >= 4
    for u in range(200):
        row = g[u][g[u] != graph_oracle.INVALID]
        assert u not in row and len(set(row)) == len(row)


def test_hnsw_cpu_recall_and_order():
    from conftest import embedding_like
    x = embedding_like(5000, dim=64, n_clusters=16)
    q = embedding_like(100, dim=64, seed=4321, n_clusters=16)
    h = O.HnswCpu(64, O.COS, 5000, threads=4)
    h.add(np.arange(5000), x)
    assert len(h) == 5000
    hk, hd = h.search(q, 10)
    tk, _, _, _ = O.exact_topk(x, q, 10, O.COS)
    assert O.recall_at_k(hk, tk) >= 0.9
    assert np.all(np.diff(hd, axis=1) >= 0)


def test_n4_fbin_ibin_roundtrip(tmp_path):
    # crates/benchmark/src/data/fbin.rs:23-148: u32 count, u32 dimension, then row-major payload
    from importlib import import_module
    ds = import_module("vector_store_b200.host.datasets")
    rng = np.random.default_rng(0)
    x = rng.standard_normal((37, 12)).astype(np.float32)
    gt = rng.integers(0, 37, size=(5, 10)).astype(np.int32)
    ds.write_fbin(str(tmp_path / "d.fbin"), x)
    ds.write_ibin(str(tmp_path / "g.ibin"), gt)
    assert ds.read_bin_header(str(tmp_path / "d.fbin")) == (37, 12)
    assert np.array_equal(ds.read_fbin(str(tmp_path / "d.fbin")), x)
    assert np.array_equal(ds.read_fbin(str(tmp_path / "d.fbin"), start=30, count=100), x[30:])
    assert np.array_equal(ds.read_ibin(str(tmp_path / "g.ibin")), gt)
    raw = open(tmp_path / "d.fbin", "rb").read()
    assert raw[:8] == np.array([37, 12], "<u4").tobytes() and len(raw) == 8 + 37 * 12 * 4


def test_golden_fixture_file_against_the_oracle():
    from golden_cases import METRIC, STORAGE, check_case, load_cases
    cases = load_cases()
    assert len(cases) >= 15
    for cid, c, keys, rows, q in cases:
        m, s = METRIC[c["metric"]], STORAGE[c["storage"]]
        kk, dd, cc, _ = O.exact_topk(rows, q[None, :], c["k"], m, s, keys=keys)
        check_case(c, kk[0][:cc[0]], dd[0][:cc[0]], O.HAMMING if s == O.B1 else m)
