"""CPU: the bench line the driver parses.  Checks the committed B200 line (profiles/r02_bench_n1.json, written by `python bench.py`
on the GPU box) against the contract of the brief: metric / unit of BASELINE.json, whole-job value, roofline against the measured
peak, CPU baseline, end-to-end leg with host buffers, clocks, launch count, and the variants the round-1 verdict asked for."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line():
    with open(os.path.join(ROOT, "profiles", "r02_bench_n1.json")) as f:
        return json.loads(f.read().strip().splitlines()[-1])


def test_headline_keys_and_baseline_metric():
    d = _line()
    with open(os.path.join(ROOT, "BASELINE.json")) as f:
        base = json.load(f)
    assert d["metric"] == "activation_stack_warps_per_sec" and d["unit"] == "warps/s"
    assert "warps" in json.dumps(base).lower()                     # BASELINE.json's metric is the activation-stack warp rate
    for k in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
              "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks", "variants"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["vs_baseline"] is None and d["dtype"] == "f32" and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["gpu_launches"] == d["steps"] > 0
    # value = units of all ranks / device time of the timed steps
    assert abs(d["value"] - d["config"]["edits_per_gpu"] / (d["ms_per_step"] * 1e-3)) <= 1e-6 * d["value"]


def test_roofline_cpu_baseline_e2e_clocks():
    d = _line()
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["kernel"] == "warp_dense_tma_kernel"
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert abs(r["achieved"] - r["algorithmic_bytes_per_launch"] / (r["ms_per_launch"] * 1e-3) / 1e9) < 1e-6 * r["achieved"]
    assert r["algorithmic_bytes_per_launch"] == d["config"]["bytes_per_warp"] * d["config"]["edits_per_gpu"]
    assert 0.5 < r["frac"] < 1.2 and "traffic_source" in r                # the sanity bound of the brief
    assert 0.9 < r["traffic"] / r["algorithmic_bytes_per_launch"] < 1.1  # DRAM traffic ~ algorithmic bytes: no wasted re-reads
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["unit"] == d["unit"] and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert 0 < e["value"] < d["value"]                                    # host buffers + PCIe inside the timed region
    k = d["clocks"]
    assert k["sm_mhz"] > 0.9 * k["sm_max_mhz"] and not set(k["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_variants_cover_the_other_baseline_configs():
    v = _line()["variants"]
    for k in ("config1", "config3", "config4_strong", "config5A", "config5B", "config5C", "set_foreground_512", "guided_step"):
        assert k in v, k
    c3 = v["config3"]
    assert c3["algorithmic_bytes"] == 62914560 and abs(c3["frac"] - c3["achieved_gbs"] / 6535.7) < 1e-3
    assert c3["ms_per_evaluation_api_direct"] < 2 * c3["ms_per_evaluation_kernels"]
    assert v["config4_strong"]["scaling"] == "strong" and v["config4_strong"]["edits_total"] == 256
    assert v["guided_step"]["bit_identical_to_eager"] is True and v["guided_step"]["fused_wall_us"] < v["guided_step"]["eager_wall_us"]
    assert v["set_foreground_512"]["converged"] is True
