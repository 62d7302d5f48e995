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


@pytest.mark.parametrize("config", ["c1", "c4"])
def test_reference_arm_line_frame_configs(config):
    """Configs 1 and 4 have no pair loop: their reference arm times the README template / calc_orient_mof over whole frames."""
    d = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--config", config], 600)
    assert BASE_KEYS <= set(d) and d["impl"] == "reference" and d["unit"] == "frames/s" and d["value"] > 1
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] == 1


@pytest.mark.gpu
def test_b200_arm_line():
    d = _run(["--steps", "3", "--warmup", "3", "--no-cpu-baseline", "--other-steps", "2"], 1500)
    assert BASE_KEYS | {"roofline", "clocks", "frames_per_s", "evaluated_pairs_per_s", "check", "other_configs", "ingest_parity"} <= set(d)
    assert d["n_gpus"] == 1 and d["steps"] == 3 and d["warmup"] >= 3 and d["scaling"] == "weak" and d["data"] == "synthetic"
    assert d["value"] > 1e12 and d["gpu_launches"] > 0 and d["ingest_parity"] == "unpinned"
    F = d["config"]["frames_per_step_per_gpu"]
    assert F * 20 >= 1000 and d["config"]["distinct_batches_per_gpu"] >= 4       # the default 20 steps cover config 2's 1000 frames
    rf = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic", "kernel", "issue_active", "inst_per_pair"} <= set(rf) and 0.3 < rf["frac"] < 1.0
    assert abs(rf["achieved"] / rf["peak"] - rf["frac"]) < 1e-9
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > F * 216000 * 12 - 1 and e["d2h_bytes_per_step"] > 0 and 0.5 * d["value"] < e["value"] <= 1.02 * d["value"]
    assert d["clocks"]["sm_mhz"] and not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    # the bench checks itself: the timed histogram equals an independent evaluation of the same frames
    assert d["check"]["parity"] == "ok" and d["check"]["frames_merged"] == 3 * F
    assert d["check"]["counted_merged"] > 0.4 * 3 * F * 72000 * 72000
    oc = d["other_configs"]
    assert sorted(oc) == ["c1", "c3", "c4", "c5"]
    for name, o in oc.items():
        assert "error" not in o, (name, o)
        assert o["check"]["parity"] == "ok", (name, o["check"])
        assert o["frames_per_s"] > 0 and o["e2e"]["frames_per_s"] > 0 and 0 < o["roofline"]["frac"] < 1.05, name
        assert o["roofline"]["bound"] == ("hbm" if name in ("c1", "c4") else "fp32_issue")
    cpp = d.get("host_cpp")
    if cpp is not None and "error" not in cpp:
        assert cpp["frames_per_s_1gpu"] > 100
