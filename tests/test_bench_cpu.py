"""CPU checks of bench.py's host logic: the reference arm's JSON line (on a shrunken gallery so that it takes a second),
the clock-sampler parser, the workload naming.  No GPU needed."""
import argparse
import importlib
import io
import json
import os
import sys
from contextlib import redirect_stdout

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture()
def bench():
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    mod = importlib.import_module("bench")
    saved = (mod.G_PER_GPU, mod.Q_TOTAL, mod.FOV, mod.SW, mod.FLOP_PER_PAIR)
    yield mod
    mod.G_PER_GPU, mod.Q_TOTAL, mod.FOV, mod.SW, mod.FLOP_PER_PAIR = saved


def test_reference_arm_line(bench, monkeypatch):
    bench.G_PER_GPU = 96                       # the real arm sweeps 10k items per query; the line's shape is what is checked
    monkeypatch.setenv("RANK", "0")
    out = io.StringIO()
    with redirect_stdout(out):
        rc = bench.run_reference(argparse.Namespace(gpus=1, steps=1, warmup=0))
    assert rc == 0
    lines = [ln for ln in out.getvalue().splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == bench.METRIC and d["unit"] == "queries/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["kind"] == "port"
    assert d["cpu_baseline"]["cores"] >= 1 and "queries" in d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["vs_baseline"] is None and d["n_gpus"] == 1 and d["steps"] == 1
    # both arms describe the workload with the same object (the driver compares them)
    assert d["config"] == bench.workload_config(1) and {"workload", "gallery_total", "gallery_per_gpu", "queries", "fov"} <= set(d["config"])


def test_cpu_sample_reports_ranks_and_ties(bench):
    """The CPU leg doubles as the bench line's parity check: ranks of the first queries and, per query, how many gallery
    items tie the threshold within fp32 round-off."""
    import numpy as np

    from oracle import witw_oracle as O
    ov, su = bench.synth_cpu(64, 8, noise=20.0, seed=3)
    qps, cores, ranks, ties, dt = bench.cpu_rank_sample(ov, su, budget_s=5.0, max_queries=8)
    assert qps > 0 and cores >= 1 and len(ranks) == len(ties) >= 2
    assert np.array_equal(ranks, O.rank_loop(ov, su, query_indices=list(range(len(ranks)))))
    assert (ties >= 0).all()


def test_reference_arm_other_ranks_stay_silent(bench, monkeypatch):
    monkeypatch.setenv("RANK", "3")
    out = io.StringIO()
    with redirect_stdout(out):
        assert bench.run_reference(argparse.Namespace(gpus=8, steps=2, warmup=1)) == 0
    assert out.getvalue() == ""


def test_clock_sampler_parses_and_skips_warmup_samples(bench):
    s = bench.ClockSampler(0)
    s.proc = type("P", (), {"terminate": lambda self: None})()
    s.lines = ["1200, 1965, 410.1, Not Active, Not Active, Not Active, Active",      # before mark(): ignored
               "1965, 1965, 880.0, Not Active, Not Active, Not Active, Not Active",
               "1950, 1965, 990.5, Not Active, Not Active, Not Active, Active",
               "garbage line", "1965, 1965, [N/A], Not Active, Active, Not Active, Not Active"]
    s.first = 1
    r = s.stop()
    assert r["sm_mhz"] == 1965.0 and r["sm_max_mhz"] == 1965.0 and r["samples"] == 3
    assert r["reasons"] == ["hw_thermal_slowdown", "sw_power_cap"]
    assert bench.ClockSampler(0).stop()["reasons"] == ["nvidia-smi unavailable"]


def test_workload_names_follow_baseline_configs(bench):
    bench.FOV, bench.Q_TOTAL, bench.G_PER_GPU = 360, 10000, 10000
    assert bench.config_name(10000) == "configs[1]"
    bench.FOV = 90
    assert bench.config_name(10000) == "configs[2]"
    bench.G_PER_GPU = 100000
    assert "configs[4]" in bench.config_name(100000)
    bench.FOV, bench.G_PER_GPU = 360, 125000
    assert bench.config_name(1000000) == "configs[3]"
    assert bench.config_name(250000).startswith("a variation")
