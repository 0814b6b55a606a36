"""The driver's bench.py contract, checked on the archived final line of the round (profiles/) and on the CPU
reference arm's static pieces: every key the task names is present with the right type, `value` is the whole-job
aggregate, the roofline / cpu_baseline / e2e objects carry their fields.  No GPU needed."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    with open(os.path.join(ROOT, "profiles", name)) as f:
        return json.loads(f.read().strip().splitlines()[-1])


def test_final_bench_line_has_the_contract_keys():
    d = _line("r01_bench_1024_t_final.json")
    for k, t in (("metric", str), ("value", float), ("unit", str), ("n_gpus", int), ("steps", int), ("warmup", int),
                 ("ms_per_step", float), ("higher_is_better", bool), ("scaling", str), ("dtype", str), ("data", str),
                 ("config", dict), ("clocks", dict), ("e2e", dict), ("gpu_launches", int), ("roofline", dict),
                 ("cpu_baseline", dict)):
        assert isinstance(d[k], t), (k, type(d[k]))
    assert d["vs_baseline"] is None and d["scaling"] == "weak" and d["higher_is_better"] and d["warmup"] >= 3
    assert "workload" in d["config"] and d["data"] == "synthetic" and d["gpu_launches"] > 0
    pairs = d["config"]["global_batch_pairs"]
    assert abs(d["value"] - pairs / (d["ms_per_step"] * 1e-3)) <= 1e-6 * d["value"]      # whole-job pairs / s
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"])
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and d["e2e"]["value"] <= d["value"] * 1.02
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and r["bound"] in ("hbm", "tensor")
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    c = d["cpu_baseline"]
    assert {"value", "unit", "cores", "kind", "sample"} <= set(c) and c["kind"] in ("reference", "port")
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_two_gpu_line_is_the_aggregate():
    d1, d2 = _line("r01_bench_1024_t_final.json"), _line("r01_bench_1024_n2_t_final.json")
    assert d2["n_gpus"] == 2 and d2["config"]["global_batch_pairs"] == 2 * d1["config"]["global_batch_pairs"]
    assert abs(d2["value"] - d2["config"]["global_batch_pairs"] / (d2["ms_per_step"] * 1e-3)) <= 1e-6 * d2["value"]
    assert d2["value"] > 1.8 * d1["value"]     # weak scaling over NVLink: > 90 % at N = 2
