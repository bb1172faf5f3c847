"""bench.py prints ONE JSON line with the keys the driver's contract names -- for the reference arm (CPU, runs
everywhere) and for the b200 arm (GPU)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches"}


def _run(*args):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], check=True, capture_output=True, text=True,
                         cwd=ROOT, timeout=900).stdout.strip().splitlines()
    assert len(out) == 1, out
    return json.loads(out[0])


def test_reference_arm_line():
    d = _run("--impl", "reference", "--steps", "1", "--warmup", "1")
    assert BASE_KEYS | {"impl", "cpu_baseline"} <= set(d)
    assert d["impl"] == "reference" and d["unit"] == "MLUPS" and d["value"] > 0 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["dtype"] == "f64" and d["data"] == "synthetic" and d["vs_baseline"] is None


@pytest.mark.gpu
def test_b200_arm_line():
    d = _run("--L", "512", "--steps", "8", "--warmup", "3")
    assert BASE_KEYS | {"roofline", "cpu_baseline", "clocks", "impl"} <= set(d)
    assert d["impl"] == "b200" and d["n_gpus"] == 1 and d["steps"] == 8 and d["warmup"] == 3 and d["higher_is_better"] is True
    assert d["gpu_launches"] == 8  # one fused kernel per step
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-3
    assert abs(r["achieved"] - 144.0 * 512 * 512 / (d["ms_per_step"] * 1e-3) / 1e9) < 0.02 * r["achieved"]
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] != d["value"]
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] > 0 and "sample" in c
    assert d["mass_drift_rel"] < 1e-12
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])


def test_c4_workload_helpers():
    """Host logic of the complete C4 workload: the step ranges between two substrate moves and the slab-wise pattern."""
    import numpy as np

    sys.path.insert(0, ROOT)
    import bench

    assert bench.segments(10, 200) == [(10, 88, True), (98, 98, True), (196, 14, False)]
    assert bench.segments(0, 98) == [(0, 98, True)] and bench.segments(97, 2) == [(97, 1, True), (98, 1, False)]
    for s0, n in [(0, 1), (5, 400), (98, 98), (195, 3)]:
        segs = bench.segments(s0, n)
        assert sum(c for _, c, _ in segs) == n and segs[0][0] == s0
        assert all(t0 + c == nxt[0] for (t0, c, _), nxt in zip(segs, segs[1:]))
        assert all(move == ((t0 + c) % bench.TMOVE == 0) for t0, c, move in segs)
    full = bench.theta_pattern(64, 48)
    parts = [bench.theta_pattern(64, 12, j0, 48) for j0 in (0, 12, 24, 36)]
    assert np.array_equal(np.concatenate(parts, axis=1), full)
    assert abs(full.mean() - 1 / 9) < 1e-12 and full.max() <= 1 / 9 + 1 / 36 + 1e-15
    h = bench.initial_height(32, workload="thermal_moving")
    assert abs(h.max() - 1.1) < 0.01 and abs(h.mean() - 1.0) < 1e-12
