"""Loader + checker for tests/golden/index_path_goldens.json (shared by the CPU-oracle and the GPU test)."""
import json
import os

import numpy as np

import oracle as O

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "index_path_goldens.json")
METRIC = {"l2sq": O.L2SQ, "cos": O.COS, "ip": O.IP}
STORAGE = {"f32": O.F32, "f16": O.F16, "bf16": O.BF16, "i8": O.I8, "b1": O.B1}


def _vec(v):
    if isinstance(v, dict):
        if "fill" in v:
            return np.full(v["dim"], v["fill"], np.float32)
        lo, hi = v["linspace"]
        return np.linspace(lo, hi, v["dim"], dtype=np.float32)
    return np.asarray(v, np.float32)


def load_cases():
    out = []
    for c in json.load(open(_PATH))["cases"]:
        keys = np.array([int(k) for k in c["corpus"]], np.uint64)
        rows = np.stack([_vec(v) for v in c["corpus"].values()])
        out.append((c["id"], c, keys, rows, _vec(c["query"])))
    return out


def check_case(c, keys, dists, metric_code):
    """keys / dists: what the implementation under test returned for the case's query (valid entries only)."""
    if "expect_keys" in c:
        assert [int(k) for k in keys[:len(c["expect_keys"])]] == c["expect_keys"], c["id"]
    if "expect_distances" in c:
        assert [float(d) for d in dists[:len(c["expect_distances"])]] == c["expect_distances"], c["id"]
    if "expect_distance_below" in c:
        assert float(dists[0]) < c["expect_distance_below"], c["id"]
    if "expect_distance_above" in c:
        assert float(dists[0]) > c["expect_distance_above"], c["id"]
    if "expect_first_set" in c:
        d0 = float(dists[0])
        assert sorted(int(k) for k, d in zip(keys, dists) if float(d) == d0) == sorted(c["expect_first_set"]), c["id"]
    if "expect_similarity" in c:
        sims = [float(O.similarity_score(d, metric_code)) for d in dists[:len(c["expect_similarity"])]]
        assert np.allclose(sims, c["expect_similarity"], atol=1e-5), c["id"]
        assert all(a > b for a, b in zip(sims, sims[1:])), c["id"]
