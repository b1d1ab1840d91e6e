"""The bench line the driver parses: the record kept under profiles/ (written by `python bench.py` on a B200, second session of round 2)
carries every key of the contract, with consistent figures."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_kept_bench_record_follows_the_contract():
    d = json.load(open(os.path.join(ROOT, "profiles", "r02b_bench_cfg3_1gpu.json")))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
              "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["unit"] == "Mrays/s" and d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["warmup"] >= 3 and d["gpu_launches"] > 0 and "workload" in d["config"] and "model" not in d["config"]
    assert d["clocks"]["reasons"] == [] and d["clocks"]["sm_mhz"] > 0.9 * d["clocks"]["sm_max_mhz"]
    e = d["e2e"]
    assert e["d2h_bytes_per_step"] == 1920 * 1080 * 16 and e["h2d_bytes_per_step"] > 0 and 0.9 < e["value"] / d["value"] <= 1.01
    r = d["roofline"]
    assert r["bound"] == "hbm" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert abs(r["achieved"] - r["algorithmic_bytes_per_launch"] / (r["avg_launch_ms"] * 1e-3) / 1e9) < 0.01 * r["achieved"]
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] > 0
    # whole-job throughput = rays / device time of the timed steps
    assert abs(d["value"] - d["rays"]["total_all_ranks"] / (d["ms_per_step"] * d["steps"]) / 1e3) < 0.01 * d["value"]
