"""Property tests of the CPU oracle (hypothesis): the checker itself is checked against plain numpy float64 on random
shapes, ragged dimensions, duplicates and tombstones.  CPU only."""
import numpy as np
from hypothesis import given, settings, strategies as st

import oracle as O


@st.composite
def case(draw):
    n = draw(st.integers(1, 60))
    dim = draw(st.integers(1, 70))
    nq = draw(st.integers(1, 5))
    k = draw(st.integers(1, 12))
    seed = draw(st.integers(0, 2 ** 31 - 1))
    dup = draw(st.booleans())
    rng = np.random.default_rng(seed)
    x = rng.integers(-8, 9, (n, dim)).astype(np.float32) / 4.0       # exactly representable: fp32 == fp64 arithmetic
    if dup and n > 3:
        x[n // 2:] = x[: n - n // 2]                                  # exact duplicates -> ties broken by key
    q = rng.integers(-8, 9, (nq, dim)).astype(np.float32) / 4.0
    keys = rng.permutation(10 * n)[:n].astype(np.uint64)
    alive = (rng.random(n) > 0.2).astype(np.uint8)
    return x, q, k, keys, alive


@settings(max_examples=60, deadline=None)
@given(case())
def test_exact_topk_l2_matches_numpy_on_exactly_representable_inputs(c):
    x, q, k, keys, alive = c
    gk, gd, gc, _ = O.exact_topk(x, q, k, O.L2SQ, O.F32, keys=keys, alive=alive)
    d = ((q[:, None, :].astype(np.float64) - x[None, :, :].astype(np.float64)) ** 2).sum(-1)
    live = np.where(alive)[0]
    for i in range(len(q)):
        order = sorted(live, key=lambda j: (d[i, j], int(keys[j])))[:k]
        assert gc[i] == len(order)
        assert [int(v) for v in gk[i, :gc[i]]] == [int(keys[j]) for j in order]
        assert np.array_equal(gd[i, :gc[i]].astype(np.float64), d[i, order])
        assert np.all(np.isinf(gd[i, gc[i]:]))                       # padding: +inf distances


@settings(max_examples=40, deadline=None)
@given(case())
def test_cosine_and_ip_are_consistent_with_their_definitions(c):
    x, q, k, keys, alive = c
    dm_ip = O.distance_matrix(x, q, O.IP, O.F32)
    dm_cos = O.distance_matrix(x, q, O.COS, O.F32)
    dot = q.astype(np.float64) @ x.astype(np.float64).T
    assert np.allclose(dm_ip, 1.0 - dot, atol=1e-4)
    qn, xn = np.linalg.norm(q.astype(np.float64), axis=1)[:, None], np.linalg.norm(x.astype(np.float64), axis=1)[None, :]
    both0, one0 = (qn == 0) & (xn == 0), (qn == 0) ^ (xn == 0)
    with np.errstate(divide="ignore", invalid="ignore"):
        want = np.clip(1.0 - dot / (qn * xn), 0.0, 2.0)
    want = np.where(both0, 0.0, np.where(one0, 1.0, want))
    assert np.allclose(dm_cos, want, atol=1e-5)
    assert dm_cos.min() >= 0.0 and dm_cos.max() <= 2.0              # the range distance.rs:66-69 requires
