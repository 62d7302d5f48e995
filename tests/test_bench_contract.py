"""bench.py prints ONE JSON line per run with the keys the driver reads; both arms are checked here (the reference arm on the
CPU, with a tiny sample; the B200 arm under the gpu marker, with a few steps)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "e2e", "gpu_launches"}


def _run(args, timeout):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    return json.loads(lines[0])


@pytest.mark.parametrize("config", ["c2", "c3"])
def test_reference_arm_line(config):
    d = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-sample", "40", "--config", config], 600)
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "rdf_pair_distances_per_s" and d["unit"] == "pair-distances/s" and d["higher_is_better"] is True
    assert d["value"] > 1e6 and d["vs_baseline"] is None and d["gpu_launches"] == 0
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert {"value", "unit", "cores", "kind", "sample"} <= set(cb) and cb["kind"] in ("reference", "port") and cb["cores"] == 1
    assert cb["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    if cb["kind"] == "reference":
        assert d["cpu_all_cores"]["processes"] == (os.cpu_count() or 1) and d["cpu_all_cores"]["value"] > 0


@pytest.mark.gpu
def test_b200_arm_line():
    d = _run(["--steps", "3", "--warmup", "3", "--no-cpu-baseline"], 900)
    assert BASE_KEYS | {"roofline", "clocks", "frames_per_s", "evaluated_pairs_per_s"} <= set(d)
    assert d["n_gpus"] == 1 and d["steps"] == 3 and d["warmup"] >= 3 and d["scaling"] == "weak" and d["data"] == "synthetic"
    assert d["value"] > 1e12 and d["gpu_launches"] > 0
    rf = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic", "kernel"} <= set(rf) and 0.3 < rf["frac"] < 1.0
    assert abs(rf["achieved"] / rf["peak"] - rf["frac"]) < 1e-9
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 8 * 216000 * 12 - 1 and e["d2h_bytes_per_step"] > 0 and 0.5 * d["value"] < e["value"] <= 1.02 * d["value"]
    assert d["clocks"]["sm_mhz"] and not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    # the histogram really was filled: 72 000^2 - 72 000 ordered pairs per frame, minus those beyond Rm
    assert d["check"]["frames_merged"] == 3 * 8 and d["check"]["counted_pairs_merged"] > 0.4 * 3 * 8 * 72000 * 72000
