"""CPU suite: the committed bench lines carry every key of the bench contract (profiles/r02_bench_*.json are the lines
`python bench.py` printed on B200s), and bench.py's static workload description is consistent with SURVEY 8(d)."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    return json.load(open(os.path.join(ROOT, "profiles", name)))


def test_one_gpu_line_has_the_contract_keys():
    l = _line("r02_bench_1gpu.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in l, k
    assert l["n_gpus"] == 1 and l["higher_is_better"] is True and l["scaling"] == "weak" and l["vs_baseline"] is None
    assert l["warmup"] >= 3 and l["gpu_launches"] > 0 and "workload" in l["config"]
    r = l["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-3
    assert r["traffic"] is None or r["traffic"] >= r["algorithmic_bytes"]
    c = l["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    e = l["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] < l["value"]
    assert not set(l["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    # whole-step fraction = algorithmic bytes / time / measured peak
    assert abs(l["frac_of_hbm_peak"] - l["algorithmic_gbs"] / r["peak"]) < 2e-3
    for k in ("train", "train_eager", "train_completion", "train_s3dis"):
        assert l[k]["samples_per_s"] > 0 and "reference file, unmodified" in l[k]["model"], k
    assert l["train"]["params_m"] == 24.02 and l["train"]["steps"] >= 100


def test_eight_gpu_line_scales():
    one, eight = _line("r02_bench_1gpu.json"), _line("r02_bench_8gpu.json")
    assert eight["n_gpus"] == 8 and eight["metric"] == one["metric"] and eight["config"]["workload"] == one["config"]["workload"]
    assert eight["value"] / one["value"] > 7.5                                  # the operators shard with no exchange
    assert eight["train"]["samples_per_s"] / one["train"]["samples_per_s"] >= 7.0      # north star: >= 7 x at 8 GPUs
    assert eight["train_nccl_syncbn"]["samples_per_s"] < eight["train"]["samples_per_s"]


def test_algorithmic_bytes_formula():
    import sys
    sys.path.insert(0, ROOT)
    from cloud_transformers_b200.hotpath import algorithmic_bytes
    N, d, F, C = 2048, 3, 4, 32 ** 3
    b = algorithmic_bytes(N, d, F, C)
    assert b["total"] == N * (24 * d + 20 * F) + 24 * F * C
