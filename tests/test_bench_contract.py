"""bench.py contract (CPU part): the reference arm prints ONE JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line_has_the_contract_keys():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--rows", "20000", "--dim", "64",
                          "--cpu-sample", "20000", "--cpu-queries", "300", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["steps"] == 2 and d["warmup"] == 1 and d["higher_is_better"] is True
    assert d["metric"] == "ann_search_qps_at_recall10_ge_0.95" and d["unit"] == "queries/s" and d["value"] > 0
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["recall_at_10"] >= 0.95
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_our_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--rows", "1000", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode != 0                      # no CPU fallback on the product path
    assert not any(ln.startswith("{") for ln in out.stdout.splitlines())


def test_committed_bench_lines_carry_the_contract_keys():
    """The bench lines committed under profiles/ (what BASELINE.md quotes) have every key the measurement contract names,
    and their roofline arithmetic is self-consistent: achieved = algorithmic bytes per launch / kernel time."""
    import json
    for name in ("r2_bench_c3_n1.json", "r2_bench_c3_n8.json", "r2_bench_c2_n1.json"):
        d = json.load(open(os.path.join(ROOT, "profiles", name)))
        for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                    "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline"):
            assert key in d, (name, key)
        assert d["metric"] == "ann_search_qps_at_recall10_ge_0.95" and d["higher_is_better"] is True
        assert "workload" in d["config"] and "model" not in d["config"]
        assert d["config"]["recall_at_10"] >= 0.95
        assert d["gpu_launches"] > 0 and d["vs_baseline"] is None
        e = d["e2e"]
        assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] <= d["value"] * 1.02
        r = d["roofline"]
        assert r["bound"] == "hbm" and r["unit"] == "GB/s"
        achieved = r["algorithmic_bytes_per_launch"] / (r["kernel_ms_per_launch"] * 1e-3) / 1e9
        assert abs(achieved - r["achieved"]) / r["achieved"] < 1e-6
        assert abs(r["achieved"] / r["peak"] - r["frac"]) < 1e-9 and 0.3 < r["frac"] < 1.1
        # the kernel cannot take longer than the step it is part of
        batches = int(d["config"]["step"].split()[0])
        assert r["kernel_ms_per_launch"] * batches <= d["ms_per_step"] * 1.001
        if d["n_gpus"] == 1:
            cb = d["cpu_baseline"]
            assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] > 0 and "sample" in cb
            assert d["build_roofline"]["tensor"]["frac"] > 0 and d["build_roofline"]["hbm"]["frac"] > 0


def test_reference_arm_under_torchrun_prints_one_line_from_rank_0():
    """The driver launches the reference arm like our arm (torchrun for N > 1): rank 0 alone runs and prints, the other
    ranks exit 0 without work."""
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "bench.py"),
                          "--impl", "reference", "--gpus", "2", "--rows", "20000", "--dim", "64", "--cpu-sample", "20000",
                          "--cpu-queries", "200", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0
