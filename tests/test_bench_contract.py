"""The bench line the driver parses: the committed line of the last GPU run (profiles/r02_bench_line.json, written by
`python bench.py` on a B200) carries every key of the contract, with consistent values.  (CPU test: it reads the committed
artefact; the bench itself needs the GPU.)"""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    with open(os.path.join(ROOT, "profiles", name)) as f:
        return json.loads(f.read())


def test_our_arm_line_has_the_contract_keys():
    d = _line("r02_bench_line.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["data"] == "synthetic" and "workload" in d["config"] and "model" not in d["config"]
    assert abs(d["value"] - 1e3 / d["ms_per_step"]) / d["value"] < 1e-3            # value == steps / time at N = 1
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] == "tensor" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-3
    assert abs(r["achieved"] - r["flops"] / (r["kernel_ms"] * 1e-3) / 1e12) / r["achieved"] < 1e-2
    c = d["cpu_baseline"]
    for k in ("value", "unit", "cores", "kind", "sample"):
        assert k in c, k
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] <= d["value"] * 1.05
    assert d["gpu_launches"] > 0 and d["config"]["ba_status"] == 0
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    for k in ("per_op_us", "in_step_us", "ref_cuda", "extra"):
        assert k in d, k


def test_reference_arm_line():
    r = _line("r02_bench_reference_line.json")
    d = _line("r02_bench_line.json")
    assert r["impl"] == "reference" and r["metric"] == d["metric"] and r["unit"] == d["unit"]
    assert r["config"]["workload"] == d["config"]["workload"]
    assert r["e2e"]["h2d_bytes_per_step"] == 0 and r["e2e"]["d2h_bytes_per_step"] == 0 and r["e2e"]["value"] == r["value"]
    assert r["cpu_baseline"]["value"] == r["value"] and r["gpu_launches"] == 0
