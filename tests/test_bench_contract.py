"""The bench.py JSON line against the driver's contract, checked on the lines recorded from the last GPU runs of the round
(profiles/r1d_bench_*.json) and, live, on the CPU reference arm's own helpers."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    return json.load(open(os.path.join(ROOT, "profiles", name)))


def _check_common(d, n_gpus):
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert "pairs/sec" in base["metric"] and d["metric"] == "preference_pairs_per_sec" and d["unit"] == "pairs/s"
    assert d["n_gpus"] == n_gpus and d["higher_is_better"] is True and d["scaling"] == "weak" and d["data"] == "synthetic"
    assert d["vs_baseline"] is None and base["published"] == {}          # no published number for this metric
    assert d["dtype"] == "bf16" and d["warmup"] >= 3 and d["steps"] >= 1
    assert abs(d["value"] - 4 * n_gpus / (d["ms_per_step"] / 1e3)) < 1e-6 * d["value"]   # whole-job pairs / device time
    assert "workload" in d["config"] and "model" not in d["config"] and "l2" in d["config"]
    e = d["e2e"]
    assert e["unit"] == "pairs/s" and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert 0 < e["value"] < d["value"]                                    # the end-to-end leg carries the copies
    assert d["gpu_launches"] > 0
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s")
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0 < r["frac"] <= 1.0
    assert r["traffic"] is None or isinstance(r["traffic"], (int, float))
    c = d["clocks"]
    assert c["sm_mhz"] and c["sm_max_mhz"] and isinstance(c["reasons"], list)
    assert not {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(c["reasons"])


def test_recorded_1gpu_line_meets_the_contract():
    d = _line("r1d_bench_7b_1gpu.json")
    _check_common(d, 1)
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["unit"] == "pairs/s" and cb["sample"]
    assert 0 < cb["value"] < d["value"]
    for a in d["roofline_attention"]:
        assert abs(a["frac"] - a["achieved"] / a["peak"]) < 1e-9 and 0 < a["frac"] < 1


def test_recorded_2gpu_line_meets_the_contract():
    d = _line("r1d_bench_7b_2gpu.json")
    _check_common(d, 2)
    assert "cpu_baseline" not in d                                         # rank 0 at N = 1 only


def test_host_threads_and_traffic_helpers():
    import sys
    sys.path.insert(0, ROOT)
    import bench
    n = bench.host_threads()
    assert 1 <= n <= (os.cpu_count() or 1) and bench.host_threads() == n   # cached
    t, detail = bench._traffic()
    assert isinstance(t, int) and t > detail["algorithmic_bytes"] > 0 and "profiles/" in detail["source"]
