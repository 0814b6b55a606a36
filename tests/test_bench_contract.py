"""The driver's bench.py contract, checked on the archived final line of the round (profiles/) and on the CPU
reference arm's static pieces: every key the task names is present with the right type, `value` is the whole-job
aggregate, the roofline / cpu_baseline / e2e objects carry their fields.  No GPU needed."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    with open(os.path.join(ROOT, "profiles", name)) as f:
        return json.loads(f.read().strip().splitlines()[-1])


import pytest


@pytest.mark.parametrize("name", ["r01_bench_1024_t_final.json", "r02_bench_1024_g_final.json", "r02_bench_1024_i_final_default_run.json"])
def test_final_bench_line_has_the_contract_keys(name):
    d = _line(name)
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


@pytest.mark.parametrize("names", [("r01_bench_1024_t_final.json", "r01_bench_1024_n2_t_final.json", 2),
                                   ("r02_bench_1024_d_final.json", "r02_bench_1024_n2_final.json", 2),
                                   ("r02_bench_1024_d_final.json", "r02_bench_1024_n8_final.json", 8)])
def test_multi_gpu_line_is_the_aggregate(names):
    d1, d2, n = _line(names[0]), _line(names[1]), names[2]
    assert d2["n_gpus"] == n and d2["config"]["global_batch_pairs"] == n * d1["config"]["global_batch_pairs"]
    assert abs(d2["value"] - d2["config"]["global_batch_pairs"] / (d2["ms_per_step"] * 1e-3)) <= 1e-6 * d2["value"]
    assert d2["value"] > 0.9 * n * d1["value"]     # weak scaling over NVLink: > 90 %
    assert d2["roofline"] is not None and d2["e2e"]["value"] > 0


def test_multi_gpu_corr_sweep_collects_per_rank_files(monkeypatch):
    """N > 1 correlation sweep of bench.py: per-rank replicas, numbers exchanged through files (no collective);
    aggregate = sum of bytes / slowest rank.  A missing rank yields None instead of hanging."""
    import bench
    calls = {"barrier": 0}
    fake = {0: [{"op": "local", "shape": [2, 128, 256, 256], "bytes": 100e6, "sec": 1e-4},
                {"op": "global", "shape": [1, 128, 128, 128], "bytes": 1e9, "sec": 5e-4}],
            1: [{"op": "local", "shape": [2, 128, 256, 256], "bytes": 100e6, "sec": 2e-4},
                {"op": "global", "shape": [1, 128, 128, 128], "bytes": 1e9, "sec": 4e-4}]}
    monkeypatch.setenv("MASTER_PORT", "45678")
    monkeypatch.setenv("TORCHELASTIC_RUN_ID", "pytest")
    barrier = lambda: calls.__setitem__("barrier", calls["barrier"] + 1)
    for rank in (1, 0):
        monkeypatch.setattr(bench, "corr_volume_points", lambda dev, hbm, r=rank: fake[r])
        out = bench.corr_volume_multi("cpu", 6500.0, rank, 2, barrier, timeout_s=5.0)
        assert (out is None) == (rank != 0)
    assert calls["barrier"] == 2 and [o["op"] for o in out] == ["local", "global"]
    assert out[0]["n_gpus"] == 2 and abs(out[0]["GBps_aggregate"] - 200e6 / 2e-4 / 1e9) < 0.1
    assert abs(out[1]["GBps_aggregate"] - 2e9 / 5e-4 / 1e9) < 0.1 and out[1]["us_max_over_ranks"] == 500.0
    # rank 1 never reports: rank 0 gives up after the timeout
    monkeypatch.setattr(bench, "corr_volume_points", lambda dev, hbm: fake[0])
    assert bench.corr_volume_multi("cpu", 6500.0, 0, 2, barrier, timeout_s=0.3) is None
    # a rank whose sweep raises still returns (and would go on to the final barrier)
    def boom(dev, hbm):
        raise RuntimeError("kernel failed")
    monkeypatch.setattr(bench, "corr_volume_points", boom)
    assert bench.corr_volume_multi("cpu", 6500.0, 1, 2, barrier, timeout_s=0.3) is None
