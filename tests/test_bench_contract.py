"""The bench line's contract (task brief, section 4), checked on the last line a B200 run of this tree printed
(profiles/r03e_bench_default.log, profiles/r03b_bench_reference.log): keys, units, and the arithmetic between them."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name, prefix):
    path = os.path.join(ROOT, "profiles", name)
    if not os.path.exists(path):
        pytest.skip(f"{name} not in the tree")
    for raw in open(path):
        if raw.startswith(prefix):
            return json.loads(raw)
    pytest.fail(f"no JSON line in {name}")


def test_default_line_carries_every_contract_key():
    d = _line("r03e_bench_default.log", '{"metric')
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert d["metric"] == "stereo_pairs_per_sec_rendered_and_tsdf_fused" and d["unit"] == "stereo-pairs/s"
    assert base["metric"].startswith("stereo-pairs/sec")
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["higher_is_better"] is True and d["scaling"] in ("weak", "strong")
    assert d["vs_baseline"] is None  # BASELINE.md publishes no number for this metric
    assert d["dtype"] == "f32" and d["data"] == "synthetic" and d["config"]["workload"].startswith("C1")
    assert "model" not in d["config"]
    # value = whole-job throughput from the device-timed region
    assert d["value"] == pytest.approx(1e3 / d["ms_per_step"], rel=2e-3)
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert e["value"] < d["value"]  # host buffers + copies inside the timed region can only cost time
    # 1600x1200: float depth + uint8 left frame (+ 16 doubles of pose) up, two uint8 frames + float depth down
    assert e["h2d_bytes_per_step"] == 1600 * 1200 * (4 + 3) + 128 and e["d2h_bytes_per_step"] == 1600 * 1200 * (3 + 3 + 4)
    assert d["gpu_launches"] > 10 * d["steps"]
    c = d["clocks"]
    assert c["sm_mhz"] > 0.9 * c["sm_max_mhz"] and not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] == "GB/s" and r["peak"] > 1000
    assert r["frac"] == pytest.approx(r["achieved"] / r["peak"], rel=1e-2)
    assert r["achieved"] == pytest.approx(r["algorithmic_bytes_per_launch"] / (r["avg_ms"] * 1e-3) / 1e9, rel=1e-2)
    assert r["traffic"] is None or r["traffic"] > 0
    # every stage's share of the serialised step is consistent with its duration x launches
    k = d["kernels"]
    assert abs(sum(v["share"] for v in k.values()) - 1.0) < 0.08
    dom = max(k, key=lambda n: k[n]["avg_ms"] * k[n]["launches"])
    assert r["kernel"] == dom
    b = d["cpu_baseline"]
    assert b["kind"] in ("port", "reference") and b["cores"] >= 1 and b["value"] > 0 and b["sample"]


def test_reference_arm_line():
    d = _line("r03b_bench_reference.log", '{"impl')
    ours = _line("r03b_bench_default.log", '{"metric')
    assert d["impl"] == "reference" and d["metric"] == ours["metric"] and d["unit"] == ours["unit"]
    assert d["higher_is_better"] == ours["higher_is_better"]
    assert set(d["config"]) == set(ours["config"])  # same keys in both arms (the driver's same_config check)
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["kind"] in ("port", "reference")
    assert d["gpu_launches"] == 0  # none of OUR kernels runs in the reference arm
